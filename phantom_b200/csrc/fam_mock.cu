// fam_mock.cu -- device program family PHX_FAMILY_MOCK: the agents the reference's OWN tests
// use to pin the step loops, so that those known-answer tests can be replayed through the
// CUDA path (tests/test_gpu_kats.py):
//   kind 0  MockAgent            /root/reference/tests/__init__.py:28-32   (no behaviour)
//   kind 1  MockStrategicAgent   /root/reference/tests/__init__.py:35-69
//             obs = [env_view.proportion_time_elapsed]; decode_action -> []; reward 0.0;
//             is_terminated == is_truncated == (current_step == agent.num_steps);
//             counts its encode / decode / reward calls
//   kind 2  EchoAgent            the message-passing agents of the routing tests:
//             tests/network/test_tracking.py:21-26  (reply value // 2 while value > 1)
//             tests/network/test_resolver.py:24-45  (request -> response(cash / 2))
//           generate_messages: sends TestMessage(seed) to every neighbour with a HIGHER slot
//           (in slot order) when seed > 0 -- this replaces the tests' hand-made n.send() calls
//   kind 3  CodecAgent           a StrategicAgent assembled from Encoder / Decoder /
//             RewardFunction objects (agents.py:199-290): obs = the agent's encoder op list
//             (ChainedEncoder / DictEncoder / EmptyEncoder / Constant, encoders.py:53-131),
//             decode = EmptyDecoder compositions (decoders.py:54-124: no messages),
//             reward = reward_functions.Constant (reward_functions.py:26-38)
// Payload types: 0 = TestMessage(value), 1 = Request(cash), 2 = Response(cash) (int payloads).
// State words:   0 encode_obs_count, 1 decode_action_count, 2 compute_reward_count,
//                3 messages handled, 4 sum of handled payload values,
//                5 echo agents: `level` (float32) = level * 0.5 + value per handled message,
//                  float32 arithmetic with one rounding per operation (numpy float32 scalars)
// agent_iparam[slot] = {agent.num_steps or -1, echo seed value, echo mode (0 halve/1 req-resp)}
#include "phx_engine_host.cuh"
#ifndef PHX_JIT_TU
#include "phx_engine_wide_host.cuh"
#endif

namespace phx {
namespace {

enum { MK_AGENT = 0, MK_STRATEGIC = 1, MK_ECHO = 2, MK_CODEC = 3 };
enum { MK_TEST_MESSAGE = 0, MK_REQUEST = 1, MK_RESPONSE = 2 };

struct MockProgram {
  // run-time specialisation (phx_jit.cuh): where this program lives and what it is called
  static constexpr const char* JIT_SOURCE = "fam_mock.cu";
  static constexpr const char* JIT_NAME = "MockProgram";
  static constexpr int PW = 1, NWORDS = 6, VW = 0, ACTCAP = 32, RESPCAP = 32, OBS_DIM = 8,
                       ACT_DIM = 1, Q1CAP = 32;
  static constexpr int RECVCAP = 32;  // max messages one agent receives in a round
  static constexpr bool BATCHED = false, HAS_PRE = false, HAS_POST = false;

  static int q1_cap(const phx_spec& s) { return s.n_agents * 4 < 32 ? 32 : 32; }
  static int32_t validate(const phx_spec& s) {
    for (int i = 0; i < s.n_agents; ++i)
      PHX_REQUIRE(s.agent_kind[i] >= MK_AGENT && s.agent_kind[i] <= MK_CODEC, PHX_ERR_INVALID,
                  "unknown mock agent kind");
    PHX_REQUIRE(s.obs_dim == 8 && s.act_dim == 1, PHX_ERR_INVALID, "mock family: obs 8, act 1");
    for (int i = 0; i < s.n_agents; ++i) {
      int total = 0;
      for (int k = 0; k < PHX_MAX_CODEC_OPS; ++k) total += (s.agent_codec_op[i][k] >> 8) & 0xFF;
      PHX_REQUIRE(total <= 8, PHX_ERR_UNSUPPORTED, "encoder composition wider than 8 floats");
    }
    return PHX_OK;
  }

  // (every callback is a template over the context type: the same text runs on the warp-tile
  // engines (Ctx) and on the 128-lane block engine (WCtx, phx_engine_wide.cuh))
  static constexpr bool WIDE_OK = true;
  template <class C, class E>
  __device__ static void act(const C& c, int* st, bool has_action, const float*, E& out) {
    const auto& sp = *c.spec;
    if (c.kind == MK_STRATEGIC || c.kind == MK_CODEC) {
      if (has_action) st[1] += 1;  // decode_action_count; EmptyDecoder compositions return []
      return;
    }
    if (c.kind == MK_ECHO) {
      const int seed = sp.agent_iparam[c.slot][1];
      if (seed <= 0) return;
      const int type = sp.agent_iparam[c.slot][2] ? MK_REQUEST : MK_TEST_MESSAGE;
      for (int r = c.next_neighbour(c.slot); r >= 0; r = c.next_neighbour(r))
        out.send(r, type, seed);
    }
  }
  template <class C>
  __device__ static void view(const C&, const int*, int*) {}
  template <class C>
  __device__ static void pre(const C&, int*) {}
  template <class C>
  __device__ static void post(const C&, int*) {}

  template <class C, class E>
  __device__ static bool handle(const C& c, int* st, const Msg& m, E& out) {
    if (c.kind != MK_ECHO) return false;  // no handler registered: ValueError (agents.py:140)
    st[3] += 1;
    st[4] += m.p[0];
    st[5] = __float_as_int(__fadd_rn(__fmul_rn(__int_as_float(st[5]), 0.5f), (float)m.p[0]));
    if (m.type == MK_TEST_MESSAGE) {  // test_tracking.py:21-26
      if (m.p[0] > 1) out.send(m.sender, MK_TEST_MESSAGE, m.p[0] / 2);
      return true;
    }
    if (m.type == MK_REQUEST) {  // test_resolver.py:31-37
      out.send(m.sender, MK_RESPONSE, m.p[0] / 2);
      return true;
    }
    return m.type == MK_RESPONSE;  // test_resolver.py:39-45: returns []
  }

  template <class C>
  __device__ static bool encode(const C& c, int* st, float* obs) {
    if (c.kind == MK_STRATEGIC) {
      st[0] += 1;
      obs[0] = c.proportion_time_elapsed();
      return true;
    }
    if (c.kind != MK_CODEC) return false;
    st[0] += 1;
    int at = 0;
    for (int k = 0; k < PHX_MAX_CODEC_OPS; ++k) {  // ChainedEncoder / DictEncoder: in order
      const int op = c.spec->codec_op[c.slot][k];
      const int len = (op >> 8) & 0xFF;
      if (len == 0) break;
      float v = c.spec->codec_val[c.slot][k];                        // Constant / EmptyEncoder
      if ((op & 0xFF) == 1) v = c.proportion_time_elapsed();         // ElapsedTime
      if ((op & 0xFF) == 2) v = (float)c.step;                       // CurrentStep
      for (int j = 0; j < len && at < OBS_DIM; ++j) obs[at++] = v;
    }
    return true;
  }
  template <class C>
  __device__ static float reward(const C& c, int* st) {
    st[2] += 1;
    return c.kind == MK_CODEC ? (float)c.spec->agent_fparam[c.slot][0] : 0.0f;  // Constant(value)
  }
  template <class C>
  __device__ static bool terminated(const C& c, const int*) {
    return c.step == c.spec->agent_iparam[c.slot][0];
  }
  template <class C>
  __device__ static bool truncated(const C& c, const int*) {
    return c.step == c.spec->agent_iparam[c.slot][0];
  }
  template <class C>
  __device__ static void reset_agent(const C&, int*) {}  // call counters survive resets
};

}  // namespace

#ifndef PHX_JIT_TU  // a specialised translation unit only needs the program above
Family* make_mock_family(const phx_spec& s) { return make_engine_family<MockProgram>(s); }
#endif

}  // namespace phx
