// A USER-WRITTEN device program (test fixture): not part of libphx.so.  It is compiled at run
// time from this file and loaded through phx_create_user (csrc/phx_user.cuh); its reference-API
// twin, written with Python handlers, is tests/user_program/auction_env.py:build_reference.
//
// Agent kinds: 0 Bidder (strategic), 1 Book.   Payload types: 0 Bid(price), 1 Ack(rank, best).
// State words: bidder 0 last_rank, 1 wins, 2 last_best;  book 0 count, 1 best, 2 total.
#include "phx_user.cuh"

namespace phx {
namespace {

struct AuctionProgram {
  // payload words, state words, view words, obs / act widths; queue bounds: messages in flight
  // per round (Q1CAP) and per agent (ACTCAP / RESPCAP sent, RECVCAP received)
  static constexpr int PW = 2, NWORDS = 3, VW = 0, OBS_DIM = 3, ACT_DIM = 1, Q1CAP = 8,
                       ACTCAP = 8, RESPCAP = 8, RECVCAP = 8;
  static constexpr bool BATCHED = false, HAS_PRE = true, HAS_POST = false;
  // Width independent (every callback is a template over the context type): env classes of
  // 9..128 agents run on the 128-lane block engine, where the book answers up to 127 bidders.
  static constexpr bool WIDE_OK = true;
  static constexpr int WIDE_ACTCAP = 1, WIDE_RESPCAP = 127;
  enum { BIDDER = 0, BOOK = 1, BID = 0, ACK = 1 };

  template <class C>
  __device__ static void view(const C&, const int*, int*) {}

  template <class C, class E>
  __device__ static void act(const C& c, int* st, bool has_action, const float* action, E& out) {
    if (c.kind != BIDDER || !has_action) return;
    const float a0 = action[0];
    if (!(fabsf(a0) <= 1048576.0f)) {
      out.fault = PHX_FAULT_INVALID_ACTION;
      return;
    }
    // decode_action: [(book, Bid(int(round(a * 100))))]
    out.send(c.spec->agent_iparam[c.slot][0], BID, __float2int_rn(__fmul_rn(a0, 100.0f)));
  }

  template <class C>
  __device__ static void pre(const C& c, int* st) {
    if (c.kind == BOOK) st[0] = st[1] = st[2] = 0;  // a fresh book every step
  }
  template <class C>
  __device__ static void post(const C&, int*) {}

  template <class C, class E>
  __device__ static bool handle(const C& c, int* st, const Msg& m, E& out) {
    if (c.kind == BOOK) {
      if (m.type != BID) return false;
      st[0] += 1;                       // arrival rank: order dependent
      st[2] += m.p[0];
      if (m.p[0] > st[1]) st[1] = m.p[0];
      out.send(m.sender, ACK, st[0], st[1]);
      return true;
    }
    if (m.type != ACK) return false;
    st[0] = m.p[0];
    st[2] = m.p[1];
    if (m.p[0] == 1) st[1] += 1;
    return true;
  }

  template <class C>
  __device__ static bool encode(const C& c, int* st, float* obs) {
    obs[0] = ratio_rn(st[0], 4.0f, 0.25f);
    obs[1] = ratio_rn(st[1], 100.0f, 1.0f / 100.0f);
    obs[2] = ratio_rn(st[2], 100.0f, 1.0f / 100.0f);
    return true;
  }
  template <class C>
  __device__ static float reward(const C&, int* st) { return ratio_rn(4 - st[0], 4.0f, 0.25f); }
  template <class C>
  __device__ static bool terminated(const C& c, const int* st) {
    return c.kind == BIDDER && st[1] >= c.spec->iparams[0];  // retires after WINS_TO_RETIRE wins
  }
  template <class C>
  __device__ static bool truncated(const C&, const int*) { return false; }
  template <class C>
  __device__ static void reset_agent(const C&, int* st) { st[0] = st[1] = st[2] = 0; }
};

}  // namespace
}  // namespace phx

PHX_USER_PROGRAM(phx::AuctionProgram)
