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
// Payload types: 0 = TestMessage(value), 1 = Request(cash), 2 = Response(cash) (int payloads).
// State words:   0 encode_obs_count, 1 decode_action_count, 2 compute_reward_count,
//                3 messages handled, 4 sum of handled payload values
// agent_iparam[slot] = {agent.num_steps or -1, echo seed value, echo mode (0 halve/1 req-resp)}
#include "phx_engine_host.cuh"

namespace phx {
namespace {

enum { MK_AGENT = 0, MK_STRATEGIC = 1, MK_ECHO = 2 };
enum { MK_TEST_MESSAGE = 0, MK_REQUEST = 1, MK_RESPONSE = 2 };

struct MockProgram {
  static constexpr int PW = 1, NWORDS = 5, VW = 0, SEGCAP = 32, OBS_DIM = 1;
  static constexpr bool BATCHED = false;

  static int32_t validate(const phx_spec& s) {
    for (int i = 0; i < s.n_agents; ++i)
      PHX_REQUIRE(s.agent_kind[i] >= MK_AGENT && s.agent_kind[i] <= MK_ECHO, PHX_ERR_INVALID,
                  "unknown mock agent kind");
    PHX_REQUIRE(s.obs_dim == 1 && s.act_dim == 1, PHX_ERR_INVALID, "mock family: obs/act dim 1");
    return PHX_OK;
  }

  template <class E>
  __device__ static void act(const Ctx& c, int* st, bool has_action, const float*, E& out) {
    const EngineSpec& sp = *c.spec;
    if (c.kind == MK_STRATEGIC) {
      if (has_action) st[1] += 1;  // decode_action_count; returns []
      return;
    }
    if (c.kind == MK_ECHO) {
      const int seed = sp.agent_iparam[c.slot][1];
      if (seed <= 0) return;
      const int type = sp.agent_iparam[c.slot][2] ? MK_REQUEST : MK_TEST_MESSAGE;
      for (int r = c.slot + 1; r < sp.n_agents; ++r)
        if (c.has_neighbour(r)) out.send(r, type, seed);
    }
  }
  __device__ static void view(const Ctx&, const int*, int*) {}
  __device__ static void pre(const Ctx&, int*) {}
  __device__ static void post(const Ctx&, int*) {}

  template <class E>
  __device__ static bool handle(const Ctx& c, int* st, const Msg& m, E& out) {
    if (c.kind != MK_ECHO) return false;  // no handler registered: ValueError (agents.py:140)
    st[3] += 1;
    st[4] += m.p[0];
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

  __device__ static bool encode(const Ctx& c, int* st, float* obs) {
    if (c.kind != MK_STRATEGIC) return false;
    st[0] += 1;
    obs[0] = c.proportion_time_elapsed();
    return true;
  }
  __device__ static float reward(const Ctx&, int* st) {
    st[2] += 1;
    return 0.0f;
  }
  __device__ static bool terminated(const Ctx& c, const int*) {
    return c.step == c.spec->agent_iparam[c.slot][0];
  }
  __device__ static bool truncated(const Ctx& c, const int*) {
    return c.step == c.spec->agent_iparam[c.slot][0];
  }
  __device__ static void reset_agent(const Ctx&, int*) {}  // call counters survive resets
};

}  // namespace

Family* make_mock_family(const phx_spec&) { return new EngineFamily<MockProgram>(); }

}  // namespace phx
