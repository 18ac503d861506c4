// phx_engine_wide.cuh -- the generic message-queue engine for env classes of 33..128 agents:
// ONE BLOCK of 128 lanes (four warps, lane == agent slot) owns one env.
//
// Same reference semantics as phx_engine.cuh (the list of /root/reference/phantom/ functions it
// replaces is at the top of that file) and the same device-program interface, with two
// differences that follow from the width:
//   * a set of agents is PHX_MASK_WORDS = 4 words.  The env-class masks live in a WideSpec in
//     global memory (17 KB: too large for the constant bank and indexed per lane anyway), the
//     per-env sets (done agents, cached rewards / observations) in shared memory, one word per
//     warp: warp w's ballot IS word w, so a set is updated without atomics.
//   * tile collectives become block collectives: a ballot / REDUX per warp, four partial
//     results in shared memory, one __syncthreads.
// The queue is the segmented queue of the tile engine (a segment per producing agent, read in
// first-arrival order, see phx_engine.cuh), with segment capacities taken from the lowered env
// class at launch (P::wide_act_cap / wide_resp_cap, or the agent's degree), so that an exchange
// that answers 120 bidders sits next to 120 agents that send one message.  A resolver round
// (wide_round): (1) the block flattens the queue into push order (a block scan of the segment
// lengths, a cooperative copy), (2) every receiver lane scans that copy and handles its batch
// (shuffle_batches: per-receiver lists from a second block scan, Fisher-Yates in place), (3) the
// next queue's segment order = the marked first-arrival positions, compacted by ballots.  Seven
// to nine block barriers per round; what dominates is a hub agent's serial batch (barrier stall).
// A program takes part by being a template over its context type (fam_mock.cu,
// fam_digital_ads.cu: `template <class C> ... const C& c`) and never touching a raw mask:
// WCtx offers has_neighbour / next_neighbour / next_of_kind like Ctx does.
// Env-level words (ENVW / env_post, see phx_engine.cuh) are kept by every lane; env_post reads
// other agents' state through a copy of all state words in shared memory.
// Not carried over: the collective resolve hook, run-time specialisation, waiting mail (a stage
// handler that does not resolve faults with PHX_FAULT_UNRESOLVED_MAIL, as on the tile engine).
#pragma once
#include <cstddef>

#include "phx_engine.cuh"

namespace phx {

constexpr int WIDE_G = PHX_MAX_AGENTS;  // lanes of a block = agent slots of ONE env
constexpr int WIDE_MW = PHX_MASK_WORDS;
static_assert(WIDE_G == 128 && WIDE_MW == 4, "the block engine is written for 4 warps");

__device__ __forceinline__ bool wbit(const uint32_t* m, int i) {
  return (m[i >> 5] >> (i & 31)) & 1u;
}
// lowest set bit above `after` (-1: from the start) or -1
__device__ inline int wnext(const uint32_t* m, int after) {
  const int i = after + 1;
  if (i >= WIDE_G) return -1;
  int w = i >> 5;
  uint32_t cur = m[w] & (0xFFFFFFFFu << (i & 31));
  while (!cur) {
    if (++w >= WIDE_MW) return -1;
    cur = m[w];
  }
  return (w << 5) + __ffs(cur) - 1;
}

// phantom.Context as the device program sees it on the block engine (cf. Ctx).
struct WCtx {
  static constexpr int MASK_WORDS = WIDE_MW;
  const WideSpec* spec;
  int slot, kind;
  int step, stage;
  uint32_t env_id, episode;
  const uint32_t* out_mask;  // [4] adjacency row of this agent (shared memory)
  const uint32_t* in_mask;   // [4] adjacency column
  const int* views;
  int view_stride;
  const int* env;
  __device__ __forceinline__ float proportion_time_elapsed() const {
    return (float)((double)step / (double)spec->num_steps);
  }
  __device__ __forceinline__ const int* view_of(int other_slot) const {
    return views + other_slot * view_stride;
  }
  __device__ __forceinline__ bool has_neighbour(int other_slot) const {
    return wbit(out_mask, other_slot);
  }
  __device__ __forceinline__ int kind_of(int other_slot) const { return spec->kind[other_slot]; }
  __device__ __forceinline__ int iparam0_of(int other_slot) const {
    return spec->agent_iparam[other_slot][0];
  }
  __device__ __forceinline__ uint32_t rand24_hi(uint32_t stream, uint32_t idx) const {
    return rng_d24_hi(spec->seed, env_id, episode, (uint32_t)step, stream, idx);
  }
  __device__ __forceinline__ int next_neighbour(int after) const { return wnext(out_mask, after); }
  __device__ __forceinline__ int next_of_kind(int k, int after) const {
    return wnext(spec->kind_mask[k].w, after);
  }
};

// The block's dynamic shared memory (WideSmem<P>, then the queue storage).  Declared at namespace
// scope so that every queue access is derived from a __shared__ symbol and compiles to LDS / STS
// (pointers kept in a struct that is indexed at run time decay to generic loads).
extern __shared__ __align__(16) unsigned char wide_raw[];

// Segmented queue of the env in shared memory; storage and capacities are set at launch.  The
// members are BYTE OFFSETS into wide_raw.
struct WQueue {
  uint32_t head;   // uint16 [total]  recv | type << 8
  uint32_t pay;    // int32  [PW][total]
  uint32_t base;   // uint16 [G + 1]
  uint32_t cnt;    // uint8  [G]
  uint32_t order;  // uint8  [G]
  uint32_t nseg;   // int32
  int total;
  __device__ __forceinline__ int base_of(int seg) const {
    return reinterpret_cast<const uint16_t*>(wide_raw + base)[seg];
  }
  __device__ __forceinline__ uint16_t& hd(int k, int seg) const {
    return reinterpret_cast<uint16_t*>(wide_raw + head)[base_of(seg) + k];
  }
  __device__ __forceinline__ int32_t& py(int w, int k, int seg) const {
    return reinterpret_cast<int32_t*>(wide_raw + pay)[w * total + base_of(seg) + k];
  }
  __device__ __forceinline__ int cap_of(int seg) const { return base_of(seg + 1) - base_of(seg); }
  __device__ __forceinline__ uint8_t& cnt_of(int seg) const { return (wide_raw + cnt)[seg]; }
  __device__ __forceinline__ uint8_t& order_at(int i) const { return (wide_raw + order)[i]; }
  __device__ __forceinline__ int32_t& nseg_ref() const {
    return *reinterpret_cast<int32_t*>(wide_raw + nseg);
  }
};

template <int PW>
struct WEmit {
  const WQueue* q;
  const WideSpec* spec;
  int slot;
  const uint32_t* out_mask;
  int n;
  uint32_t fault;
  __device__ __forceinline__ void send(int recv, int type, int p0, int p1 = 0) {
    if (fault) return;
    if (!(spec->flags & PHX_FLAG_IGNORE_CONNECTION_ERRORS) && !wbit(out_mask, recv)) {
      fault = PHX_FAULT_NO_EDGE;  // network.py:246-250
      return;
    }
    if (!(spec->flags & PHX_FLAG_NO_PAYLOAD_CHECKS)) {  // network.py:297-331
      if (!wbit(spec->sender_ok[type].w, slot) || !wbit(spec->receiver_ok[type].w, recv)) {
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
    if (PW > 1) q->py(PW > 1 ? 1 : 0, n, slot) = p1;
    ++n;
  }
};

// Segment capacities of one agent for the acting phase / a response round.
template <class P, class = void>
struct HasWideCaps : std::false_type {};
template <class P>
struct HasWideCaps<P, std::void_t<decltype(&P::wide_act_cap), decltype(&P::wide_resp_cap)>>
    : std::true_type {};
// ... or just the bounds, when they differ from the narrow engines' ACTCAP / RESPCAP (a hub that
// answers every neighbour): static constexpr int WIDE_ACTCAP, WIDE_RESPCAP.
template <class P, class = void>
struct WideCapConst {
  static constexpr int act = P::ACTCAP, resp = P::RESPCAP;
};
template <class P>
struct WideCapConst<P, std::void_t<decltype(P::WIDE_ACTCAP), decltype(P::WIDE_RESPCAP)>> {
  static constexpr int act = P::WIDE_ACTCAP, resp = P::WIDE_RESPCAP;
};
template <class P>
__host__ __device__ inline int wide_cap(bool acting, int kind, int degree, int n_agents) {
  if constexpr (HasWideCaps<P>::value) {
    return acting ? P::wide_act_cap(kind, degree, n_agents) : P::wide_resp_cap(kind, degree, n_agents);
  } else {
    // one message per neighbour and round, up to the program's bound
    const int cap = acting ? WideCapConst<P>::act : WideCapConst<P>::resp;
    const int want = degree > 0 ? degree : 1;
    return want < cap ? want : cap;
  }
}

struct WideLayout {  // byte offsets into the dynamic shared memory of a block
  int act_total, resp_total;
  int off_pay[3], off_head[3];
  int off_flat;  // uint32 [max(act_total, resp_total)]: the round's queue in push order
  int off_list;  // uint16 [max(act_total, resp_total)]: shuffle_batches: the receivers' batch lists
  int off_mark;  // uint8  [max(act_total, resp_total)]: first-arrival marks of a round
  int bytes;
};
__host__ __device__ inline WideLayout wide_layout_rt(int pw, int act_total, int resp_total) {
  WideLayout l;
  l.act_total = act_total;
  l.resp_total = resp_total;
  int at = 0;
  for (int q = 0; q < 3; ++q) {
    l.off_pay[q] = at;
    at += 4 * pw * (q == 0 ? act_total : resp_total);
  }
  for (int q = 0; q < 3; ++q) {
    l.off_head[q] = at;
    at += 2 * (q == 0 ? act_total : resp_total);
    at = (at + 3) & ~3;
  }
  l.off_flat = at;
  at += 4 * (act_total > resp_total ? act_total : resp_total);
  l.off_list = at;
  at += (2 * (act_total > resp_total ? act_total : resp_total) + 3) & ~3;
  l.off_mark = at;
  at += ((act_total > resp_total ? act_total : resp_total) + 3) & ~3;
  l.bytes = at;
  return l;
}

template <class P>
struct WideArgs {
  const WideSpec* spec;  // global memory
  int32_t T;
  WideLayout lay;
  int4* hdr;             // [E]
  uint32_t* term;        // [E][4] PhantomEnv._terminations as a bitmask over agent slots
  uint32_t* trunc;       // [E][4]
  int32_t* state;        // [NWORDS][E][G]
  float* reward_cache;   // [E][G]
  uint32_t* reward_none; // [E][4]
  float* obs_cache;      // [E][G][O]
  uint32_t* obs_cached;  // [E][4]
  int32_t* env_state;    // [ENVW][E] env-level words of the program (nullptr if ENVW == 0)
  uint32_t* adj_env;     // [E][G][4] StochasticNetwork: per-env adjacency rows (else nullptr)
  const uint2* base_conn;
  int32_t n_base;
  StepIO io;
  FaultSink faults;
  TraceSink trace;
};

// Fixed part of the block's shared memory (the queue storage follows, WideLayout).
template <class P>
struct WideSmem {
  uint32_t adj_out[WIDE_G][WIDE_MW];
  uint32_t adj_in[WIDE_G][WIDE_MW];
  int32_t first_idx[WIDE_G];
  int32_t views[WIDE_G][P::VW > 0 ? P::VW : 1];
  uint16_t qbase[3][WIDE_G + 2];
  uint8_t qcnt[3][WIDE_G];
  uint8_t qorder[3][WIDE_G];
  int32_t qnseg[3];
  uint32_t term[WIDE_MW], trunc[WIDE_MW], rnone[WIDE_MW], ocached[WIDE_MW];
  int32_t red[WIDE_MW];
  uint32_t redu[WIDE_MW];
  int32_t bcast;
  int32_t mark_part[2][WIDE_MW];  // compaction of the first-arrival marks (wide_round)
  int16_t pbase[WIDE_G];  // flattening a round's queue: first position of the i-th segment visited
  uint8_t pseg[WIDE_G];   // ... and its producer
  // programs with env-level words: every agent's state, published before env_post
  int32_t pub[EnvWords<P>::value > 0 ? (P::NWORDS > 0 ? P::NWORDS : 1) : 1][EnvWords<P>::value > 0 ? WIDE_G : 1];
};

__device__ __forceinline__ int wide_sum(int v, int32_t* red) {
  v = __reduce_add_sync(0xFFFFFFFFu, v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const int s = red[0] + red[1] + red[2] + red[3];
  __syncthreads();  // `red` may be rewritten right away
  return s;
}
__device__ __forceinline__ uint32_t wide_min(uint32_t v, uint32_t* red) {
  v = __reduce_min_sync(0xFFFFFFFFu, v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t s = min(min(red[0], red[1]), min(red[2], red[3]));
  __syncthreads();
  return s;
}
// exclusive prefix sum of v over the block's lanes (slot order) and the block total
__device__ __forceinline__ int wide_excl_scan(int v, int32_t* red, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += u;
  }
  if (lane == 31) red[warp] = incl;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int w = 0; w < WIDE_MW; ++w)
    if (w < warp) woff += red[w];
  total = red[0] + red[1] + red[2] + red[3];
  __syncthreads();  // `red` may be rewritten right away
  return woff + incl - v;
}
__device__ __forceinline__ int wide_popc(const uint32_t* m) {
  return __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
}

// StochasticNetwork.resample_connectivity for one agent's row (cf. resample_adj_row).
__device__ inline void wide_resample_row(uint64_t seed, const uint2* base, int n_base,
                                         uint32_t env_id, uint32_t episode, int slot,
                                         uint32_t* row) {
  uint32_t r[WIDE_MW] = {0u, 0u, 0u, 0u};
  for (int c = 0; c < n_base; ++c) {
    const uint2 bc = base[c];
    const int u = bc.x & 0xFF, v = (bc.x >> 8) & 0xFF;
    if (u != slot && v != slot) continue;
    if (base_connection_exists(seed, env_id, episode, c, bc.y)) {
      const int o = u == slot ? v : u;
#pragma unroll
      for (int w = 0; w < WIDE_MW; ++w)
        if (w == (o >> 5)) r[w] |= 1u << (o & 31);
    }
  }
#pragma unroll
  for (int w = 0; w < WIDE_MW; ++w) row[w] = r[w];
}

// One resolver round (resolvers.py:137-158): the block flattens the current queue into global
// push order (segments in first-arrival order of their producers); every receiver lane scans
// that copy, handles the messages addressed to it and appends its responses to its segment of
// `qn`; `qn`'s segment order = the receivers by the position of their first message.  Returns
// the responses pushed.
template <class P, bool TRACK>
__device__ __forceinline__ int wide_round(const WideArgs<P>& a, const WCtx& ctx, int* st,
                                          bool has_ctx, const WQueue& qc, const WQueue& qn, WideSmem<P>& sm,
                                          int round, uint32_t& fault_key, int& traced, size_t row,
                                          bool trace_lane, uint32_t flat_off, uint32_t list_off,
                                          uint32_t mark_off, int& k_batch) {
  constexpr int INF = 0x7FFFFFFF;
  const int slot = ctx.slot;
  WEmit<P::PW> resp{&qn, ctx.spec, slot, ctx.out_mask, 0, 0u};
  int first = INF;
  bool bad_type = false;
  const int nseg = qc.nseg_ref();

  // ---- 1. flatten the queue into push order: flat[p] = recv | type << 8 | producer << 16 |
  // index in the producer's segment << 24.  (Walking the segments per receiver costs four
  // DEPENDENT shared-memory loads per segment -- order, count, base, head -- in every lane; the
  // flat copy is one broadcast load per message.)  Lane i takes the i-th segment visited; a block
  // exclusive scan of the segment lengths gives its first position.
  int c_i = 0, seg_i = 0;
  if (slot < nseg) {
    seg_i = qc.order_at(slot);
    c_i = qc.cnt_of(seg_i);
  }
  int tot = 0;
  const int pb = wide_excl_scan(c_i, sm.red, tot);
  sm.pbase[slot] = (int16_t)(slot < nseg ? pb : 0x7FFF);
  sm.pseg[slot] = (uint8_t)seg_i;
  __syncthreads();
  uint32_t* flat = reinterpret_cast<uint32_t*>(wide_raw + flat_off);
  for (int p = slot; p < tot; p += WIDE_G) {
    // the segment that holds position p: the LAST i with pbase[i] <= p (empty segments share
    // their successor's first position)
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sm.pbase[mid] <= p) lo = mid;
      else hi = mid - 1;
    }
    const int seg = sm.pseg[lo], k = p - sm.pbase[lo];
    flat[p] = (uint32_t)qc.hd(k, seg) | ((uint32_t)seg << 16) | ((uint32_t)k << 24);
    (wide_raw + mark_off)[p] = 0;  // first-arrival marks of this round (step 3)
  }
  __syncthreads();

  // ---- 2. every receiver handles its batch, in push order
  if constexpr (P::BATCHED) {
    if (has_ctx) P::batch_begin(ctx, st);
  }
  auto deliver = [&](const uint32_t f) {
    const int seg = (int)((f >> 16) & 0xFFu), k = (int)(f >> 24);
    Msg m;
    m.sender = seg;
    m.type = (int)((f >> 8) & 0xFFu);
    m.p[0] = qc.py(0, k, seg);
    m.p[1] = P::PW > 1 ? qc.py(P::PW > 1 ? 1 : 0, k, seg) : 0;
    if (!P::handle(ctx, st, m, resp)) bad_type = true;  // agents.py:140-143
  };
  // (TRACK = the FULL build of the kernel: message tracking and / or shuffle_batches; the lean
  // build carries neither -- the shuffle path inlines the program's handlers a second time and
  // cost the 122-agent market 33 % when it sat in the same kernel)
  if (!TRACK || !(ctx.spec->flags & PHX_FLAG_SHUFFLE_BATCHES)) {
    for (int pos = 0; pos < tot; ++pos) {
      const uint32_t f = flat[pos];
      if ((int)(f & 0xFFu) != slot) continue;
      if (first == INF) first = pos;  // first-arrival position of this receiver
      if (!has_ctx) continue;         // done agent: mail dropped (resolvers.py:143-144)
      if (!wbit(ctx.in_mask, (int)((f >> 16) & 0xFFu))) continue;  // delivery-time edge filter (:146-148)
      deliver(f);
    }
  } else {
    // BatchResolver(shuffle_batches=True), resolvers.py:146-151: the batch is first reduced to the
    // messages whose edge still exists, then shuffled (the contract's Fisher-Yates, see
    // shuffle_batch in phx_engine.cuh), then handled.  The receivers' lists share one array of
    // `tot` entries: count, block scan for the offsets, fill, shuffle in place.
    int n_mine = 0;
    for (int pos = 0; pos < tot; ++pos) {
      const uint32_t f = flat[pos];
      if ((int)(f & 0xFFu) != slot) continue;
      if (first == INF) first = pos;
      n_mine += has_ctx && wbit(ctx.in_mask, (int)((f >> 16) & 0xFFu));
    }
    int all = 0;
    const int off = wide_excl_scan(n_mine, sm.red, all);
    uint16_t* list = reinterpret_cast<uint16_t*>(wide_raw + list_off) + off;
    if (n_mine > 0) {
      int j = 0;
      for (int pos = first; j < n_mine; ++pos) {
        const uint32_t f = flat[pos];
        if ((int)(f & 0xFFu) == slot && wbit(ctx.in_mask, (int)((f >> 16) & 0xFFu)))
          list[j++] = (uint16_t)pos;
      }
      shuffle_list(ctx.spec->seed, ctx.env_id, ctx.episode, ctx.step, slot, list, n_mine, k_batch);
      for (j = 0; j < n_mine; ++j) deliver(flat[list[j]]);
    }
  }
  if constexpr (P::BATCHED) {
    if (has_ctx && first != INF) P::batch_end(ctx, st, resp);
  }
  // mark the position of this receiver's first message (a byte per position, cleared by the
  // flattening pass; this lane is its only writer): the marked positions, compacted, are the
  // next queue's segment order
  uint8_t* marks = wide_raw + mark_off;
  if (first != INF) marks[first] = 1;
  if (bad_type)
    fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) |
                                   PHX_FAULT_UNKNOWN_MSG_TYPE);
  if (resp.fault)
    fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) | resp.fault);
  qn.cnt_of(slot) = (uint8_t)resp.n;
  const int total_next = wide_sum(resp.n, sm.red);  // (barrier: the marks and qn are visible)

  // next queue's segment order = the receivers by the position of their first message
  // (resolvers.py:126,142: dict insertion order): compact the marked positions -- a ballot per
  // warp, the four warp totals through shared memory (two buffers, one barrier per 128 positions)
  int nrecv = 0;
  {
    const int lane = slot & 31, warp = slot >> 5;
    for (int c0 = 0; c0 < tot; c0 += WIDE_G) {
      const int p = c0 + slot;
      const uint32_t f = p < tot ? flat[p] : 0u;
      const bool mark = p < tot && marks[p] != 0;
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, mark);
      int32_t* part = sm.mark_part[(c0 / WIDE_G) & 1];
      if (lane == 0) part[warp] = __popc(b);
      __syncthreads();
      int woff = 0;
#pragma unroll
      for (int w = 0; w < WIDE_MW; ++w)
        if (w < warp) woff += part[w];
      if (mark) qn.order_at(nrecv + woff + __popc(b & ((1u << lane) - 1u))) = (uint8_t)(f & 0xFFu);
      nrecv += part[0] + part[1] + part[2] + part[3];
    }
  }
  if (slot == 0) qn.nseg_ref() = nrecv;
  __syncthreads();
  if (TRACK && trace_lane) {
    for (int si = 0; si < nrecv; ++si) {
      const int seg = qn.order_at(si);
      for (int k = 0; k < qn.cnt_of(seg); ++k) {
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

// Everything a block sets up once per launch (step and reset kernels).
template <class P>
struct WideBlock {
  WideSmem<P>* sm;
  WQueue q[3];  // 0 acting phase, 1 / 2 response rounds
  uint32_t flat_off, list_off, mark_off;
};

template <class P>
__device__ __forceinline__ void wide_setup(const WideArgs<P>& a, int slot, int e, WideBlock<P>& wb) {
  const WideSpec& sp = *a.spec;
  WideSmem<P>& sm = *reinterpret_cast<WideSmem<P>*>(wide_raw);
  const uint32_t dyn = (uint32_t)((sizeof(WideSmem<P>) + 15) & ~(size_t)15);
  wb.sm = &sm;
  wb.flat_off = dyn + (uint32_t)a.lay.off_flat;
  wb.list_off = dyn + (uint32_t)a.lay.off_list;
  wb.mark_off = dyn + (uint32_t)a.lay.off_mark;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    wb.q[q].pay = dyn + (uint32_t)a.lay.off_pay[q];
    wb.q[q].head = dyn + (uint32_t)a.lay.off_head[q];
    wb.q[q].base = (uint32_t)(offsetof(WideSmem<P>, qbase) + sizeof(sm.qbase[0]) * q);
    wb.q[q].cnt = (uint32_t)(offsetof(WideSmem<P>, qcnt) + sizeof(sm.qcnt[0]) * q);
    wb.q[q].order = (uint32_t)(offsetof(WideSmem<P>, qorder) + sizeof(sm.qorder[0]) * q);
    wb.q[q].nseg = (uint32_t)(offsetof(WideSmem<P>, qnseg) + sizeof(int32_t) * q);
    wb.q[q].total = q == 0 ? a.lay.act_total : a.lay.resp_total;
  }
  const bool is_agent = slot < sp.n_agents;
  // adjacency rows of this env
#pragma unroll
  for (int w = 0; w < WIDE_MW; ++w)
    sm.adj_out[slot][w] = !is_agent ? 0u
                          : a.adj_env ? a.adj_env[((size_t)e * WIDE_G + slot) * WIDE_MW + w]
                                      : sp.adj[slot].w[w];
  // segment capacities from the lowered env class: caps into first_idx, then prefix sums
  const int deg = is_agent ? __popc(sp.adj[slot].w[0]) + __popc(sp.adj[slot].w[1]) +
                                 __popc(sp.adj[slot].w[2]) + __popc(sp.adj[slot].w[3])
                           : 0;
  const int kind = is_agent ? sp.kind[slot] : -1;
  for (int phase = 0; phase < 2; ++phase) {
    sm.first_idx[slot] = is_agent ? wide_cap<P>(phase == 0, kind, deg, sp.n_agents) : 0;
    __syncthreads();
    int base = 0;
    for (int j = 0; j < slot; ++j) base += sm.first_idx[j];
    if (phase == 0) {
      sm.qbase[0][slot] = (uint16_t)base;
      if (slot == WIDE_G - 1) sm.qbase[0][WIDE_G] = (uint16_t)(base + sm.first_idx[slot]);
    } else {
      sm.qbase[1][slot] = sm.qbase[2][slot] = (uint16_t)base;
      if (slot == WIDE_G - 1)
        sm.qbase[1][WIDE_G] = sm.qbase[2][WIDE_G] = (uint16_t)(base + sm.first_idx[slot]);
    }
    __syncthreads();
  }
}

// in-row of `slot` from the out-rows in shared memory (after a barrier)
template <class P>
__device__ __forceinline__ void wide_in_row(WideSmem<P>& sm, int n_agents, int slot) {
  uint32_t in[WIDE_MW] = {0u, 0u, 0u, 0u};
  for (int s = 0; s < n_agents; ++s) {
    const uint32_t bit = (sm.adj_out[s][slot >> 5] >> (slot & 31)) & 1u;
#pragma unroll
    for (int w = 0; w < WIDE_MW; ++w)
      if (w == (s >> 5)) in[w] |= bit << (s & 31);
  }
#pragma unroll
  for (int w = 0; w < WIDE_MW; ++w) sm.adj_in[slot][w] = in[w];
}

template <class P, bool TRACK>
__device__ __forceinline__ void wide_step_body(const WideArgs<P>& a) {
  const WideSpec& sp = *a.spec;
  const int slot = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x;
  const bool is_agent = slot < sp.n_agents;
  WideBlock<P> wb;
  wide_setup<P>(a, slot, e, wb);
  WideSmem<P>& sm = *wb.sm;
  __syncthreads();
  wide_in_row<P>(sm, sp.n_agents, slot);
  if (slot < WIDE_MW) {
    sm.term[slot] = a.term[(size_t)e * WIDE_MW + slot];
    sm.trunc[slot] = a.trunc[(size_t)e * WIDE_MW + slot];
    sm.rnone[slot] = sp.env_kind != PHX_ENV_BASE ? a.reward_none[(size_t)e * WIDE_MW + slot] : 0u;
    sm.ocached[slot] = sp.env_kind == PHX_ENV_FSM ? a.obs_cached[(size_t)e * WIDE_MW + slot] : 0u;
  }
  __syncthreads();

  const int kind = is_agent ? sp.kind[slot] : -1;
  const int sidx = is_agent ? sp.sidx[slot] : -1;
  const bool strategic = sidx >= 0;
  const int S = sp.n_strategic, O = sp.obs_dim;
  const bool cached_env = sp.env_kind != PHX_ENV_BASE;

  int4 h = a.hdr[e];
  int st[P::NWORDS > 0 ? P::NWORDS : 1];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) st[w] = a.state[((size_t)w * sp.E + e) * WIDE_G + slot];
  float rcache = cached_env ? a.reward_cache[(size_t)e * WIDE_G + slot] : 0.f;
  uint32_t fault_key = 0xFFFFFFFFu;  // (phase << 16 | slot << 8 | code), smallest wins
  // env-level words: every lane keeps the same copy
  constexpr int EW = EnvWords<P>::value;
  int envw[EW > 0 ? EW : 1] = {0}, envsnap[EW > 0 ? EW : 1] = {0};
  if constexpr (EW > 0) {
#pragma unroll
    for (int w = 0; w < EW; ++w) envw[w] = a.env_state[(size_t)w * sp.E + e];
  }

  WCtx ctx;
  ctx.spec = &sp;
  ctx.slot = slot;
  ctx.kind = kind;
  ctx.env = envsnap;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.views = &sm.views[0][0];
  ctx.view_stride = P::VW > 0 ? P::VW : 1;
  ctx.out_mask = sm.adj_out[slot];
  ctx.in_mask = sm.adj_in[slot];

  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * sp.E + e;
    float act[P::ACT_DIM];
    bool has_action_now = true;
#pragma unroll
    for (int j = 0; j < P::ACT_DIM; ++j) act[j] = 0.f;
    if (strategic) {
      const size_t arow = row * S + sidx;
#pragma unroll
      for (int j = 0; j < P::ACT_DIM; ++j) act[j] = a.io.actions[arow * P::ACT_DIM + j];
      if (a.io.action_mask) has_action_now = a.io.action_mask[arow] != 0;
    }
    h.x += 1;  // env.py:252
    ctx.step = h.x;
    ctx.episode = (uint32_t)h.y;
    ctx.stage = h.z;
    const bool was_done = is_agent && (wbit(sm.term, slot) || wbit(sm.trunc, slot));
    const bool has_ctx = is_agent && !was_done;  // env.py:344-348: no context for done agents
    if constexpr (EW > 0) {  // the EnvView of this step (env.py:340)
#pragma unroll
      for (int w = 0; w < EW; ++w) envsnap[w] = envw[w];
    }

    if (P::VW > 0) {  // start-of-step snapshot of every agent's public state
      if (is_agent) P::view(ctx, st, &sm.views[slot][0]);
      __syncthreads();
    }

    // ---- who acts / observes / is rewarded (env.py:320-336; fsm.py:276-320;
    // stackelberg.py:133-140): this lane's bit of each set
    bool acting = true, observing = strategic, rewarded = strategic;
    int next_stage = h.z;
    bool handled = false, resolves = true;
    if (sp.env_kind == PHX_ENV_FSM) {
      acting = wbit(sp.stage_acting[h.z].w, slot);
      next_stage = sp.stage_next[h.z];
      handled = sp.stage_rule[h.z][SR_HANDLER] != 0;
      resolves = !handled || sp.stage_rule[h.z][SR_RESOLVES] != 0;
      if (!sp.stage_rewarded_none[h.z]) {  // fsm.py:315-320
        rewarded = wbit(sp.stage_rewarded[h.z].w, slot);
        observing = wbit(sp.stage_acting[next_stage].w, slot);
      }
    } else if (sp.env_kind == PHX_ENV_STACKELBERG) {
      const bool leaders_turn = (h.x & 1) == 1;
      const bool lead = wbit(sp.leaders.w, slot), follow = wbit(sp.followers.w, slot);
      acting = leaders_turn ? lead : follow;
      observing = leaders_turn ? follow : lead;
      rewarded = acting;
    }

    // ---- acting phase
    WEmit<P::PW> out{&wb.q[0], &sp, slot, ctx.out_mask, 0, 0u};
    if (has_ctx && acting) P::act(ctx, st, strategic && has_action_now, act, out);
    wb.q[0].cnt_of(slot) = (uint8_t)out.n;
    // the segments are visited in the order the agents acted: slot order, or the stage's own
    // list (fsm.py:276-277, stackelberg.py:133-140)
    int n_acting_segs = sp.n_agents;
    if (!sp.any_act_order) {
      wb.q[0].order_at(slot) = (uint8_t)slot;
    } else {
      const int phase = sp.env_kind == PHX_ENV_FSM ? ctx.stage
                        : sp.env_kind == PHX_ENV_STACKELBERG ? ((h.x & 1) == 1 ? 0 : 1) : 0;
      n_acting_segs = sp.n_act[phase];
      if (slot < n_acting_segs) wb.q[0].order_at(slot) = (uint8_t)sp.act_order[phase][slot];
    }
    if (slot == 0) wb.q[0].nseg_ref() = n_acting_segs;
    if (out.fault) fault_key = min(fault_key, (0u << 16) | ((uint32_t)slot << 8) | out.fault);
    int pending = wide_sum(out.n, sm.red);

    int traced = 0;
    const bool trace_lane = TRACK && slot == 0 && a.trace.rows != nullptr;
    if (trace_lane) {  // pushes of the acting phase, in global push order
      for (int oi = 0; oi < n_acting_segs; ++oi)
        for (int si = wb.q[0].order_at(oi), k = 0; k < wb.q[0].cnt_of(si); ++k) {
          if (traced < a.trace.cap)
            a.trace.rows[row * a.trace.cap + traced] =
                make_int4((int)(((uint32_t)wb.q[0].hd(k, si) << 8) | (uint32_t)si),
                          wb.q[0].py(0, k, si), P::PW > 1 ? wb.q[0].py(P::PW > 1 ? 1 : 0, k, si) : 0, 0);
          ++traced;
        }
    }

    if (!resolves && pending > 0) {  // the mail would wait for a later step's resolve
      fault_key = min(fault_key, (1u << 16) | (0xFFu << 8) | PHX_FAULT_UNRESOLVED_MAIL);
      pending = 0;
    }
    if (has_ctx && resolves) P::pre(ctx, st);  // env.py:170-173

    // ---- BatchResolver.resolve (resolvers.py:128-163)
    int k_batch = 0;  // batches this receiver has shuffled in this step (shuffle_batches only)
    for (int round = 0; pending > 0; ++round) {
      if (sp.round_limit >= 0 && round >= sp.round_limit) {  // resolvers.py:160-163
        fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | (0xFFu << 8) | PHX_FAULT_ROUND_LIMIT);
        break;
      }
      // (selected by value: a run-time index into wb.q[] would put the queues in local memory)
      const WQueue qc = round == 0 ? wb.q[0] : ((round - 1) & 1) ? wb.q[2] : wb.q[1];
      const WQueue qn = (round & 1) ? wb.q[2] : wb.q[1];
      pending = wide_round<P, TRACK>(a, ctx, st, has_ctx, qc, qn, sm, round, fault_key, traced, row,
                                     trace_lane, wb.flat_off, wb.list_off, wb.mark_off, k_batch);
    }
    if (trace_lane) a.trace.cnt[row] = traced;

    if (has_ctx && resolves) P::post(ctx, st);  // env.py:175-178
    if constexpr (EW > 0) {  // the env class's own post_message_resolution override
      if (resolves) {        // (uniform over the block: the stage is an env-level word)
#pragma unroll
        for (int w = 0; w < P::NWORDS; ++w) sm.pub[w][slot] = st[w];
        __syncthreads();
        P::env_post(ctx, envw, [&](int s_, auto w_) { return sm.pub[decltype(w_)::value][s_]; });
        __syncthreads();
      }
    }

    // ---- the stage's env handler picks the next stage (fsm.py:294-307)
    if (handled) {
      next_stage = stage_rule_pick(sp, h.z, [&](int okind, int rs, int rw, int constant) {
        if (okind == PHX_RULE_STEP) return (int)h.x;
        if (okind == PHX_RULE_ENV_WORD) {
          int v = constant;
          if constexpr (EW > 0) {
#pragma unroll
            for (int w = 0; w < EW; ++w)
              if (w == rw) v = envw[w];
          }
          return v;
        }
        if (okind != PHX_RULE_AGENT_WORD) return constant;
        if (slot == rs) {
          int mine = 0;
#pragma unroll
          for (int w = 0; w < P::NWORDS; ++w)
            if (w == rw) mine = st[w];
          sm.bcast = mine;
        }
        __syncthreads();
        const int v = sm.bcast;
        __syncthreads();
        return v;
      });
      if (!((sp.stage_allowed[h.z] >> next_stage) & 1u)) {
        fault_key = min(fault_key, (0xFFFEu << 16) | (0xFFu << 8) | PHX_FAULT_BAD_TRANSITION);
        next_stage = h.z;
      }
      if (!sp.stage_rewarded_none[h.z]) observing = wbit(sp.stage_acting[next_stage].w, slot);
    }

    // ---- outputs for strategic agents (env.py:273-303; fsm.py:322-378;
    // stackelberg.py:149-194)
    bool obs_now = false, rew_now = false;
    float obs_val[P::OBS_DIM] = {};
    float rew_val = 0.f;
    bool t_flag = false, u_flag = false;
    if (strategic && has_ctx) {
      if (observing) obs_now = P::encode(ctx, st, obs_val);  // None -> false
      if (sp.env_kind == PHX_ENV_BASE) {
        if (obs_now) {  // env.py:281-284: reward only travels with an observation
          rew_val = P::reward(ctx, st);
          rew_now = true;
        }
      } else if (rewarded) {
        rew_val = P::reward(ctx, st);
        rew_now = true;
        rcache = rew_val;
      }
      t_flag = P::terminated(ctx, st);
      u_flag = P::truncated(ctx, st);
    }
    {  // the env's sets: warp w's ballot is word w
      const uint32_t tb = __ballot_sync(0xFFFFFFFFu, t_flag), ub = __ballot_sync(0xFFFFFFFFu, u_flag);
      const uint32_t ob = __ballot_sync(0xFFFFFFFFu, obs_now), rb = __ballot_sync(0xFFFFFFFFu, rew_now);
      if (lane == 0) {
        sm.term[warp] |= tb;
        sm.trunc[warp] |= ub;
        if (cached_env) {
          sm.rnone[warp] &= ~rb;  // _rewards.update(rewards)
          if (sp.env_kind == PHX_ENV_FSM) sm.ocached[warp] |= ob;
        }
      }
    }
    __syncthreads();
    const bool all_term = wide_popc(sm.term) == S;                             // env.py:308-310
    const bool all_trunc = (h.x == sp.num_steps) || wide_popc(sm.trunc) == S;  // env.py:312-318
    const bool terminal = all_term || all_trunc;
    if (sp.env_kind == PHX_ENV_FSM) h.z = next_stage;  // fsm.py:355
    const bool my_rnone = wbit(sm.rnone, slot), my_ocached = wbit(sm.ocached, slot);

    if (strategic) {
      const size_t orow = row * S + sidx;
      uint8_t om = 0, rm = 0;
      float r_out = 0.f;
      if (sp.env_kind == PHX_ENV_BASE) {
        om = obs_now;
        rm = rew_now;
        r_out = rew_val;
      } else if (sp.env_kind == PHX_ENV_FSM) {
        float* oc = a.obs_cache + ((size_t)e * WIDE_G + slot) * O;
        if (obs_now) {
#pragma unroll
          for (int j = 0; j < P::OBS_DIM; ++j)
            if (j < O) oc[j] = obs_val[j];
        }
        if (terminal) {  // fsm.py:360-375: flush the caches
          om = my_ocached;
          if (om && !obs_now) {
#pragma unroll
            for (int j = 0; j < P::OBS_DIM; ++j)
              if (j < O) obs_val[j] = oc[j];
          }
          rm = my_rnone ? 2 : 1;
          r_out = rcache;
        } else {  // fsm.py:378: last computed reward of every agent observing now
          om = obs_now;
          if (obs_now) {
            rm = my_rnone ? 2 : 1;
            r_out = rcache;
          }
        }
      } else {  // Stackelberg
        om = obs_now;
        if (terminal) {  // stackelberg.py:180-187: the whole reward cache
          rm = my_rnone ? 2 : 1;
          r_out = rcache;
        } else if (obs_now && !my_rnone) {  // stackelberg.py:190-194
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
    if (slot == 0 && a.io.all_done)
      reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(all_term, all_trunc);
    __syncthreads();  // every lane has read the sets of this step

    // ---- PHX_FLAG_AUTO_RESET: the step that ends the episode also resets the env
    if ((sp.flags & PHX_FLAG_AUTO_RESET) && terminal) {
      h.x = 0;
      h.y += 1;
      h.z = sp.initial_stage;
      ctx.step = 0;
      ctx.episode = (uint32_t)h.y;
      ctx.stage = h.z;
      if (slot < WIDE_MW) {
        sm.term[slot] = sm.trunc[slot] = 0u;
        sm.rnone[slot] = cached_env ? sp.strategic_mask.w[slot] : 0u;
      }
      if (a.adj_env) {  // Network.reset of a StochasticNetwork resamples first (network.py:450-453)
        if (is_agent)
          wide_resample_row(sp.seed, a.base_conn, a.n_base, ctx.env_id, ctx.episode, slot,
                            sm.adj_out[slot]);
        __syncthreads();
        wide_in_row<P>(sm, sp.n_agents, slot);
      }
      if (is_agent) P::reset_agent(ctx, st);
      __syncthreads();
      if constexpr (EW > 0) {  // reset() builds a fresh EnvView (fsm.py:232-236)
#pragma unroll
        for (int w = 0; w < EW; ++w) envsnap[w] = envw[w];
      }
      if (P::VW > 0) {
        if (is_agent) P::view(ctx, st, &sm.views[slot][0]);
        __syncthreads();
      }
      bool first_obs = strategic;
      if (sp.env_kind == PHX_ENV_FSM) first_obs = first_obs && wbit(sp.stage_acting[sp.initial_stage].w, slot);
      if (sp.env_kind == PHX_ENV_STACKELBERG) first_obs = first_obs && wbit(sp.leaders.w, slot);
      if (strategic) {
        const size_t orow = row * S + sidx;
        bool got = false;
        if (first_obs) got = P::encode(ctx, st, obs_val);
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
  __syncthreads();
  if (slot == 0) {
    a.hdr[e] = h;
    if constexpr (EW > 0) {
#pragma unroll
      for (int w = 0; w < EW; ++w) a.env_state[(size_t)w * sp.E + e] = envw[w];
    }
  }
  if (slot < WIDE_MW) {
    a.term[(size_t)e * WIDE_MW + slot] = sm.term[slot];
    a.trunc[(size_t)e * WIDE_MW + slot] = sm.trunc[slot];
    if (cached_env) {
      a.reward_none[(size_t)e * WIDE_MW + slot] = sm.rnone[slot];
      if (sp.env_kind == PHX_ENV_FSM) a.obs_cached[(size_t)e * WIDE_MW + slot] = sm.ocached[slot];
    }
  }
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * WIDE_G + slot] = st[w];
  if (cached_env) a.reward_cache[(size_t)e * WIDE_G + slot] = rcache;
  if (a.adj_env) {
#pragma unroll
    for (int w = 0; w < WIDE_MW; ++w)
      a.adj_env[((size_t)e * WIDE_G + slot) * WIDE_MW + w] = sm.adj_out[slot][w];
  }
  // first fault of the env in event order (phase, then agent order)
  const uint32_t fk = wide_min(fault_key, sm.redu);
  if (slot == 0 && fk != 0xFFFFFFFFu) raise_fault(a.faults, e, fk & 0xFFu);
}

template <class P, bool TRACK>
__global__ void __launch_bounds__(WIDE_G) wide_step_kernel(const WideArgs<P> a) {
  wide_step_body<P, TRACK>(a);
}

// PhantomEnv.reset / FiniteStateMachineEnv.reset / StackelbergEnv.reset for masked envs.
template <class P>
__device__ __forceinline__ void wide_reset_body(const WideArgs<P>& a, const uint8_t* env_mask,
                                                float* obs, uint8_t* obs_mask, bool agents_only) {
  const WideSpec& sp = *a.spec;
  const int slot = threadIdx.x;
  const int e = blockIdx.x;
  if (env_mask != nullptr && env_mask[e] == 0) return;  // (block-uniform)
  const bool is_agent = slot < sp.n_agents;
  WideBlock<P> wb;
  wide_setup<P>(a, slot, e, wb);
  WideSmem<P>& sm = *wb.sm;
  int4 h = a.hdr[e];
  int st[P::NWORDS > 0 ? P::NWORDS : 1];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) st[w] = a.state[((size_t)w * sp.E + e) * WIDE_G + slot];
  h.x = 0;
  h.y += 1;
  h.z = sp.env_kind == PHX_ENV_FSM ? sp.initial_stage : 0;
  constexpr int EW = EnvWords<P>::value;
  int envw[EW > 0 ? EW : 1] = {0};  // env-level words survive a reset unless the program says otherwise
  if constexpr (EW > 0) {
#pragma unroll
    for (int w = 0; w < EW; ++w) envw[w] = a.env_state[(size_t)w * sp.E + e];
  }
  WCtx ctx;
  ctx.spec = &sp;
  ctx.slot = slot;
  ctx.kind = is_agent ? sp.kind[slot] : -1;
  ctx.env = envw;
  ctx.step = 0;
  ctx.stage = h.z;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.episode = (uint32_t)h.y;
  ctx.views = &sm.views[0][0];
  ctx.view_stride = P::VW > 0 ? P::VW : 1;
  ctx.out_mask = sm.adj_out[slot];
  ctx.in_mask = sm.adj_in[slot];
  if (a.adj_env) {
    // StochasticNetwork.reset resamples the edges before the agents reset (network.py:450-453);
    // the constructor's own sample (add_connection, :389-391) is the draw of "episode -1"
    if (is_agent)
      wide_resample_row(sp.seed, a.base_conn, a.n_base, ctx.env_id,
                        agents_only ? 0xFFFFFFFFu : ctx.episode, slot, sm.adj_out[slot]);
#pragma unroll
    for (int w = 0; w < WIDE_MW; ++w)
      a.adj_env[((size_t)e * WIDE_G + slot) * WIDE_MW + w] = is_agent ? sm.adj_out[slot][w] : 0u;
  }
  __syncthreads();
  wide_in_row<P>(sm, sp.n_agents, slot);
  __syncthreads();
  if (is_agent) P::reset_agent(ctx, st);  // Network.reset -> agent.reset() (network.py:179-184)
  if (agents_only) {  // PhantomEnv.__init__ ends with agent.reset() only (env.py:122-124)
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * WIDE_G + slot] = st[w];
    return;
  }
  if (P::VW > 0) {
    if (is_agent) P::view(ctx, st, &sm.views[slot][0]);
    __syncthreads();
  }
  const int sidx = is_agent ? sp.sidx[slot] : -1;
  bool first_obs = sidx >= 0;
  if (sp.env_kind == PHX_ENV_FSM) first_obs = first_obs && wbit(sp.stage_acting[sp.initial_stage].w, slot);
  if (sp.env_kind == PHX_ENV_STACKELBERG) first_obs = first_obs && wbit(sp.leaders.w, slot);
  if (sidx >= 0) {
    float obs_val[P::OBS_DIM] = {};
    bool got = false;
    if (first_obs) got = P::encode(ctx, st, obs_val);
    const size_t orow = (size_t)e * sp.n_strategic + sidx;
    if (obs && got) {
#pragma unroll
      for (int j = 0; j < P::OBS_DIM; ++j)
        if (j < sp.obs_dim) obs[orow * sp.obs_dim + j] = obs_val[j];
    }
    if (obs_mask) obs_mask[orow] = got;
  }
  if (slot == 0) a.hdr[e] = h;
  if (slot < WIDE_MW) {
    a.term[(size_t)e * WIDE_MW + slot] = 0u;
    a.trunc[(size_t)e * WIDE_MW + slot] = 0u;
    if (sp.env_kind != PHX_ENV_BASE)  // _rewards = None
      a.reward_none[(size_t)e * WIDE_MW + slot] = sp.strategic_mask.w[slot];
  }
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * WIDE_G + slot] = st[w];
}

template <class P>
__global__ void __launch_bounds__(WIDE_G)
wide_reset_kernel(const WideArgs<P> a, const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                  bool agents_only) {
  wide_reset_body<P>(a, env_mask, obs, obs_mask, agents_only);
}

// A program says that its callbacks are width independent with `static constexpr bool WIDE_OK`.
template <class P, class = void>
struct IsWideOk : std::false_type {};
template <class P>
struct IsWideOk<P, std::void_t<decltype(P::WIDE_OK)>> : std::bool_constant<P::WIDE_OK> {};

}  // namespace phx
