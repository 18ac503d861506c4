/* phx.h -- C ABI of libphx: the B200-native batched env-step engine.
 *
 * This is the drop-in boundary for the data-parallel hot path of jpmorganchase/Phantom
 * (reference @ 9ce42f2, v2.2.0).  The reference has no FFI: its boundary is the Python
 * plugin API, and one call of `PhantomEnv.step` advances ONE env object
 * (phantom/env.py:239-303).  Each entry point below replaces that call (and the Python
 * list-of-envs loops around it, phantom/utils/rllib/rollout.py:289-363) for a whole batch
 * of independent env instances whose state lives in HBM.
 *
 * Conventions
 *   - plain C, no torch types; every buffer is a raw pointer + an implied shape.
 *   - "device pointer" = CUDA device memory on the handle's device, contiguous, aligned to
 *     16 bytes; the caller owns every I/O buffer (a torch tensor's data_ptr()).
 *   - E = num_envs of the handle, S = spec.n_strategic, A = spec.act_dim, O = spec.obs_dim.
 *   - all work is ordered on `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises the host except the *_host variants, phx_get_field, phx_poll_errors,
 *     phx_get_trace and phx_destroy.
 *   - every int-returning function returns PHX_OK (0) or a negative phx_status; the
 *     message for the last failure on the calling thread is phx_last_error().
 *   - device-side faults (the reference's exceptions raised from inside step()) never abort
 *     a kernel: they are recorded in a sticky per-env error word (first code wins) and
 *     collected with phx_poll_errors().
 *   - a handle is not thread-safe; one handle per device.
 */
#ifndef PHX_H_
#define PHX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHX_ABI_VERSION 7

#define PHX_MAX_AGENTS 128  /* agent slots per env                                   */
#define PHX_MAX_TYPES 16    /* payload types per env class                           */
#define PHX_MAX_STAGES 8    /* FSM stages                                            */
#define PHX_MASK_WORDS 4    /* PHX_MAX_AGENTS / 32                                   */
#define PHX_MAX_PARAMS 16
#define PHX_TRACE_WORDS 4   /* one traced message = 4 x int32 (see phx_get_trace)    */
#define PHX_MAX_CODEC_OPS 6 /* encoder ops per agent (Chained/Dict encoder composition)*/
#define PHX_MAX_BASE_CONNECTIONS 528 /* StochasticNetwork base connections (32*33/2)     */

typedef enum phx_status {
  PHX_OK = 0,
  PHX_ERR_INVALID = -1,      /* bad argument / spec rejected by the family lowering   */
  PHX_ERR_CUDA = -2,         /* a CUDA runtime call failed                            */
  PHX_ERR_UNSUPPORTED = -3,  /* valid in the reference, no device program for it      */
  PHX_ERR_NO_DEVICE = -4
} phx_status;

/* Device fault codes; each mirrors the exception the reference raises from step(). */
typedef enum phx_fault {
  PHX_FAULT_NONE = 0,
  PHX_FAULT_NO_EDGE = 1,          /* NetworkError, phantom/network.py:246-249          */
  PHX_FAULT_BAD_PAYLOAD_TYPE = 2, /* NetworkError, phantom/network.py:315-331          */
  PHX_FAULT_UNKNOWN_MSG_TYPE = 3, /* ValueError,   phantom/agents.py:140-143           */
  PHX_FAULT_ROUND_LIMIT = 4,      /* RuntimeError, phantom/resolvers.py:160-163        */
  PHX_FAULT_BAD_TRANSITION = 5,   /* FSMRuntimeError, phantom/fsm.py:304-307           */
  PHX_FAULT_QUEUE_OVERFLOW = 6,   /* engine capacity exceeded (no reference analogue)  */
  PHX_FAULT_INVALID_ACTION = 7,   /* non-finite / out-of-contract action value
                                     (ValueError/OverflowError from int(round(a)))     */
  PHX_FAULT_UNRESOLVED_MAIL = 8,  /* a stage handler that does not call resolve_network()
                                     left messages queued: the reference carries them into
                                     a later step's resolve; so does the thread-per-env
                                     engine (<= 8 agents), the tile and block engines
                                     refuse with this fault                              */
  PHX_FAULT_PLAN_MISMATCH = 9     /* a step kernel specialised with a STATIC message schedule
                                     (phx_jit_source) saw a send its device program did not
                                     declare: the program's act_sends / handle_sends signature
                                     is wrong (no reference analogue; use the generic kernel) */
} phx_fault;

/* Which step loop drives the env (phantom/env.py, fsm.py, stackelberg.py). */
typedef enum phx_env_kind {
  PHX_ENV_BASE = 0,
  PHX_ENV_FSM = 1,
  PHX_ENV_STACKELBERG = 2
} phx_env_kind;

/* Device program families (one fused step kernel each). */
typedef enum phx_family {
  PHX_FAMILY_SUPPLY_CHAIN = 1, /* examples/environments/supply_chain/supply_chain.py  */
  PHX_FAMILY_MOCK = 2,         /* the agents of the reference's own step-loop KATs
                                  (tests/__init__.py:28-69)                           */
  PHX_FAMILY_MARKET = 3,       /* 3-stage FSM market (BASELINE config C3)             */
  PHX_FAMILY_STACKELBERG = 4,  /* leader/follower pricing game (BASELINE config C4)   */
  PHX_FAMILY_DENSE = 5,        /* dense-graph broadcast + batch aggregation (C5)      */
  PHX_FAMILY_SIMPLE_MARKET = 7,/* examples/environments/simple_market/ (2-stage FSM, env-level
                                  post_message_resolution + custom EnvView field)         */
  PHX_FAMILY_DIGITAL_ADS = 8,  /* examples/environments/digital_ads_market/ (auction in a
                                  handle_batch override, three response rounds)           */
  PHX_FAMILY_SUPPLY_CHAIN2 = 6,/* multi-shop supply chain with agent supertypes
                                  (docs/user/tutorial2.rst)                           */
  PHX_FAMILY_USER = 100        /* a device program the CALLER compiled (phx_create_user):
                                  the reference's "bring your own agent classes"
                                  (phantom/agents.py:48-60) without rebuilding libphx */
} phx_family;

typedef enum phx_exec_mode {
  PHX_EXEC_AUTO = 0,   /* fastest kernel that is valid for the spec                   */
  PHX_EXEC_QUEUE = 1,  /* force the generic engine with a tile of lanes per env       */
  PHX_EXEC_FAST = 2,   /* force the schedule-specialised kernel; create fails if the
                          spec is outside its domain                                  */
  PHX_EXEC_THREAD = 3, /* force the generic engine with one thread per env (<= 8 agents) */
  PHX_EXEC_WIDE = 4    /* force the generic engine with one 128-lane block per env (chosen
                          automatically for 33..128 agents; programs that support it)   */
} phx_exec_mode;

enum {
  PHX_FLAG_IGNORE_CONNECTION_ERRORS = 1 << 0, /* Network(ignore_connection_errors=True)   */
  PHX_FLAG_NO_PAYLOAD_CHECKS = 1 << 1,        /* Network(enforce_msg_payload_checks=False)*/
  PHX_FLAG_TRACK_MESSAGES = 1 << 2,           /* BatchResolver(enable_tracking=True)      */
  PHX_FLAG_AUTO_RESET = 1 << 3,               /* reset an env in the step that ends its
                                                 episode; obs then holds the reset obs   */
  PHX_FLAG_STOCHASTIC_NETWORK = 1 << 4,       /* StochasticNetwork: every env resamples its
                                                 edges from base_* at reset
                                                 (phantom/network.py:439-453); <= 32 agents */
  PHX_FLAG_SHUFFLE_BATCHES = 1 << 5           /* BatchResolver(shuffle_batches=True),
                                                 phantom/resolvers.py:150-151              */
};

/* The device form of an FSM stage's env handler (phantom/fsm.py:294-307).  The reference calls an
 * arbitrary Python function that may run self.resolve_network() and returns the next stage;
 * the device evaluates the declarative equivalent
 *     [self.resolve_network()]; return then_stage if <lhs> <cmp> rhs else else_stage
 * after the stage's message resolution, on the env's state at that point. */
typedef enum phx_rule_lhs {
  PHX_RULE_ALWAYS = 0,     /* unconditional: returns then_stage                            */
  PHX_RULE_STEP = 1,       /* env.current_step (already incremented, phantom/fsm.py:266)   */
  PHX_RULE_AGENT_WORD = 2, /* int32 state word `word` of agent slot `slot`                 */
  PHX_RULE_ENV_WORD = 3,   /* int32 env-level word `word` (families with env-level state)  */
  PHX_RULE_CONST = 4       /* right-hand sides only: the int32 constant `rhs`              */
} phx_rule_lhs;
typedef enum phx_cmp { PHX_CMP_LT = 0, PHX_CMP_LE, PHX_CMP_EQ, PHX_CMP_NE, PHX_CMP_GE, PHX_CMP_GT } phx_cmp;
/* OR-ed into phx_rule_term.cmp: both operands are float32 values (a float32 state word, or the
 * bits of a float32 constant in `rhs`) and are compared as such */
#define PHX_CMP_F32 8

/* handler == 2: the handler is an if / elif / ... / else chain.  Branch b holds when ALL of its
 * n_terms comparisons `<lhs operand> <cmp> <rhs operand>` hold (an operand is a phx_rule_lhs kind
 * + slot / word, the rhs may also be a constant); the first branch that holds returns its `then`
 * stage, none -> rule_else.  (OR = two branches with the same `then`.) */
#define PHX_RULE_BRANCHES 4
#define PHX_RULE_TERMS 2
typedef struct phx_rule_term {
  int32_t lhs, slot, word;             /* left operand: phx_rule_lhs STEP / AGENT_WORD / ENV_WORD  */
  int32_t cmp;                         /* phx_cmp                                                  */
  int32_t rhs_kind, rhs_slot, rhs_word;/* right operand: phx_rule_lhs STEP / ..._WORD / CONST      */
  int32_t rhs;                         /* the constant of PHX_RULE_CONST                           */
} phx_rule_term;
typedef struct phx_rule_branch {
  int32_t n_terms;                     /* 1 .. PHX_RULE_TERMS                                      */
  int32_t then;                        /* stage index returned when every term holds               */
  phx_rule_term term[PHX_RULE_TERMS];
} phx_rule_branch;

typedef struct phx_stage {
  uint32_t acting[PHX_MASK_WORDS];   /* FSMStage.acting_agents as a slot bitmask         */
  uint32_t rewarded[PHX_MASK_WORDS]; /* FSMStage.rewarded_agents                         */
  int32_t rewarded_is_none;          /* rewarded_agents is None (phantom/fsm.py:315-317) */
  int32_t next_stage;                /* next_stages[0]: the next stage of a handler-less
                                        stage (phantom/fsm.py:284-292)                   */
  int32_t handler;                   /* 0 = no env handler; 1 = the single rule_* comparison
                                        below; 2 = the rule_branch chain
                                        (phantom/fsm.py:294-302)                         */
  uint32_t next_allowed;             /* FSMStage.next_stages as a stage bitmask; a handler
                                        returning a stage outside it faults with
                                        PHX_FAULT_BAD_TRANSITION (phantom/fsm.py:304-307) */
  int32_t rule_resolves;             /* the handler calls self.resolve_network() (a stage
                                        WITH a handler is not resolved otherwise,
                                        phantom/fsm.py:280-283)                          */
  int32_t rule_lhs, rule_slot, rule_word, rule_cmp, rule_rhs; /* phx_rule_lhs, phx_cmp   */
  int32_t rule_then, rule_else;      /* stage indices                                    */
  int32_t rule_n_branches;           /* handler == 2: 1 .. PHX_RULE_BRANCHES             */
  phx_rule_branch rule_branch[PHX_RULE_BRANCHES];
  /* The ORDER in which the stage's acting agents act.  The reference hands the user's list to
   * _handle_acting_agents (FSMStage.acting_agents, phantom/fsm.py:276-277; StackelbergEnv's
   * leader_agents / follower_agents, phantom/stackelberg.py:133-140), so the push order of a
   * step's messages follows that list, not the network's agent order.  0 = slot order (the
   * list is ascending in slot); else the n_act_order slots of `acting` in list order.
   * PHX_ENV_STACKELBERG: stages[0] / stages[1] carry the leaders' / followers' order. */
  int32_t n_act_order;
  uint8_t act_order[PHX_MAX_AGENTS];
} phx_stage;

/* Flat description of one env class, lowered from the Python objects
 * (PhantomEnv / Network / Agent / FSMStage ...) by phantom_b200/lowering.py. */
typedef struct phx_spec {
  uint32_t struct_size; /* sizeof(phx_spec), ABI guard */
  int32_t family;       /* phx_family       */
  int32_t env_kind;     /* phx_env_kind     */
  int32_t exec_mode;    /* phx_exec_mode    */
  uint32_t flags;       /* PHX_FLAG_*       */
  int32_t num_steps;    /* PhantomEnv.num_steps */
  int32_t round_limit;  /* BatchResolver.round_limit, -1 = None */
  int32_t trace_capacity; /* max traced messages per env per step (tracking only) */

  int32_t n_agents;                       /* agent slots, in Network.agents order      */
  int32_t n_strategic;                    /* number of StrategicAgent slots            */
  int32_t agent_kind[PHX_MAX_AGENTS];     /* family-specific device-program id per slot*/
  int32_t strategic_index[PHX_MAX_AGENTS];/* slot -> index among strategic agents, -1  */
  uint32_t adjacency[PHX_MAX_AGENTS][PHX_MASK_WORDS]; /* bit r of row s: edge s -> r   */

  int32_t n_payload_types;
  uint32_t type_sender_ok[PHX_MAX_TYPES][PHX_MASK_WORDS];   /* slots allowed to send   */
  uint32_t type_receiver_ok[PHX_MAX_TYPES][PHX_MASK_WORDS]; /* slots allowed to receive*/

  int32_t n_stages;      /* PHX_ENV_FSM */
  int32_t initial_stage;
  phx_stage stages[PHX_MAX_STAGES];

  uint32_t leaders[PHX_MASK_WORDS];   /* PHX_ENV_STACKELBERG */
  uint32_t followers[PHX_MASK_WORDS];

  int32_t obs_dim; /* floats per strategic agent in the obs tensors (family maximum) */
  int32_t act_dim; /* floats per strategic agent in the action tensors               */

  int32_t iparams[PHX_MAX_PARAMS]; /* family parameters (see the family header)      */
  double fparams[PHX_MAX_PARAMS];
  int32_t agent_iparam[PHX_MAX_AGENTS][4]; /* per-slot family parameters, e.g. the slot
                                              of the peer an agent addresses           */
  double agent_fparam[PHX_MAX_AGENTS][4];
  /* Encoder composition (phantom/encoders.py:64-131) lowered to a per-agent op list:
   * op = opcode | length << 8 (opcode 0 constant, 1 proportion_time_elapsed, 2 current_step),
   * evaluated in order into the agent's obs row; 0 terminates the list. */
  int32_t agent_codec_op[PHX_MAX_AGENTS][PHX_MAX_CODEC_OPS];
  float agent_codec_val[PHX_MAX_AGENTS][PHX_MAX_CODEC_OPS];

  /* PHX_FLAG_STOCHASTIC_NETWORK: StochasticNetwork._base_connections in insertion order
   * (phantom/network.py:379-399).  At every reset connection c of env e exists iff
   * uniform01(stream 5, step 0, idx c) < base_rate[c]  (np.random.random() < rate). */
  int32_t n_base_connections;
  uint8_t base_u[PHX_MAX_BASE_CONNECTIONS];
  uint8_t base_v[PHX_MAX_BASE_CONNECTIONS];
  double base_rate[PHX_MAX_BASE_CONNECTIONS];
} phx_spec;

typedef struct phx_env phx_env; /* opaque */

/* State columns readable through phx_get_field / writable through phx_set_field.
 * Generic ids below; family columns start at PHX_FIELD_FAMILY (see family headers). */
typedef enum phx_field {
  PHX_FIELD_STEP = 0,       /* int32 [E]   PhantomEnv.current_step                   */
  PHX_FIELD_EPISODE = 1,    /* int32 [E]   resets seen - 1 (RNG contract coordinate) */
  PHX_FIELD_STAGE = 2,      /* int32 [E]   FSM current stage index                   */
  PHX_FIELD_TERMINATED = 3, /* uint32[E,W] PhantomEnv._terminations as a bitmask over agent
                               slots (only strategic slots are ever set), W words   */
  PHX_FIELD_TRUNCATED = 4,  /* uint32[E,W]                                           */
  PHX_FIELD_ERROR = 5,      /* uint32[E]   sticky fault word                         */
  PHX_FIELD_ADJACENCY = 6,  /* uint32[E,n_agents] per-env adjacency rows of a
                               StochasticNetwork (bit r of row s: edge s -> r)       */
  PHX_FIELD_ENV_STATE = 7,  /* int32 [E]   env-level state word `index` of env classes that keep
                               state of their own (e.g. simple_market's avg_price, float64
                               bits in words 0/1)                                        */
  PHX_FIELD_FAMILY = 16
} phx_field;

int32_t phx_abi_version(void);
uint32_t phx_sizeof_spec(void); /* sizeof(phx_spec) as compiled: binding self-check */
const char* phx_last_error(void);
int32_t phx_device_count(void);

/* Create the device-resident batch.  Replaces `env_class(**env_config)` executed once per
 * env replica (phantom/utils/rllib/train.py:185-187, rollout.py:289-291).
 *   seed       base seed of the RNG contract (oracle/rng.py)
 *   env_offset global index of this handle's env 0 (multi-GPU sharding: results do not
 *              depend on how envs are split over handles)                             */
int32_t phx_create(const phx_spec* spec, int32_t num_envs, int32_t device, uint64_t seed,
                   int64_t env_offset, phx_env** out);
/* The same for an env class whose device program is NOT part of libphx (spec->family ==
 * PHX_FAMILY_USER): `cubin_path` names a cubin compiled for sm_100a from a translation unit that
 * ends with PHX_USER_PROGRAM(Prog) (phantom_b200/csrc/phx_user.cuh; compile with
 * `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -I <csrc>`).  The program
 * supplies the callbacks of the agent classes (decode_action / generate_messages, message
 * handlers, hooks, encode_observation, compute_reward, is_terminated / is_truncated, reset);
 * libphx supplies the step loop, the state in HBM and every other entry point of this header.
 * Env classes of at most 8 agents (thread-per-env engine). */
int32_t phx_create_user(const phx_spec* spec, const char* cubin_path, int32_t num_envs,
                        int32_t device, uint64_t seed, int64_t env_offset, phx_env** out);
void phx_destroy(phx_env* env);

int32_t phx_num_envs(const phx_env* env);
/* name of the kernel variant chosen by phx_create ("fast" / "queue<G>") */
const char* phx_exec_name(const phx_env* env);

/* PhantomEnv.reset (phantom/env.py:185-237; fsm.py:195-251; stackelberg.py:53-109) for
 * every env (env_mask == NULL) or for the envs whose mask byte is non-zero.
 *   env_mask  device uint8 [E] or NULL
 *   obs       device float [E,S,O]   initial observations (rows of unmasked envs untouched)
 *   obs_mask  device uint8 [E,S]     1 where the reference's reset() returned an obs    */
int32_t phx_reset(phx_env* env, const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                  void* stream);

/* PhantomEnv.step / FiniteStateMachineEnv.step / StackelbergEnv.step for every env.
 *   actions     device float [E,S,A]
 *   action_mask device uint8 [E,S] or NULL (= all present); 0 means the agent is absent
 *               from the `actions` mapping and falls back to generate_messages()
 *               (phantom/env.py:330-333)
 *   obs         device float [E,S,O]; rows with obs_mask == 0 are left untouched
 *   obs_mask    device uint8 [E,S]   agent is a key of Step.observations
 *   reward      device float [E,S]   float32(Step.rewards[aid])
 *   reward_mask device uint8 [E,S]   0 = not a key of Step.rewards, 1 = value present,
 *                                    2 = key present with value None (phantom/fsm.py:378)
 *   term, trunc device uint8 [E,S]   Step.terminations / truncations; 255 = key absent
 *                                    (agent was already done)
 *   all_done    device uint8 [E,2]   terminations["__all__"], truncations["__all__"]    */
int32_t phx_step(phx_env* env, const float* actions, const uint8_t* action_mask, float* obs,
                 uint8_t* obs_mask, float* reward, uint8_t* reward_mask, uint8_t* term,
                 uint8_t* trunc, uint8_t* all_done, void* stream);

/* T consecutive steps in ONE launch; env state stays on chip between steps.  All arrays
 * get a leading T: actions [T,E,S,A], obs [T,E,S,O], ...  Replaces the inner loop of
 * phantom/utils/rllib/rollout.py:321-363 when actions are known up front (open-loop
 * evaluation, replay) and is what bench.py times as the fused step kernel.              */
int32_t phx_rollout(phx_env* env, int32_t num_steps_T, const float* actions,
                    const uint8_t* action_mask, float* obs, uint8_t* obs_mask, float* reward,
                    uint8_t* reward_mask, uint8_t* term, uint8_t* trunc, uint8_t* all_done,
                    void* stream);

/* Same as phx_rollout but every buffer is HOST memory (ideally from phx_host_alloc, i.e.
 * pinned): copies the actions to the device, runs, copies all outputs back, and
 * synchronises.  This is the end-to-end call bench.py reports as `e2e`.                 */
int32_t phx_rollout_host(phx_env* env, int32_t num_steps_T, const float* actions,
                         const uint8_t* action_mask, float* obs, uint8_t* obs_mask,
                         float* reward, uint8_t* reward_mask, uint8_t* term, uint8_t* trunc,
                         uint8_t* all_done);

void* phx_host_alloc(uint64_t bytes); /* pinned host memory (cudaMallocHost) */
void phx_host_free(void* p);

/* Copy one state column to / from HOST memory (metrics, parity tests; replaces reading
 * live Python agent attributes, phantom/metrics.py:230-231).  `index` selects the agent
 * instance for per-agent family columns.  Synchronises.                                 */
int32_t phx_get_field(phx_env* env, int32_t field, int32_t index, void* host_out,
                      uint64_t out_bytes);
int32_t phx_set_field(phx_env* env, int32_t field, int32_t index, const void* host_in,
                      uint64_t in_bytes);

/* Reduce one int32 state column over the envs ON THE DEVICE (metrics: replaces a Python loop
 * over env objects reading `agent.<attr>`, phantom/metrics.py:230-231,348-352).  The column
 * `field` is viewed as rows of `width` int32 words; word `col` of every row is reduced.
 * Returns the exact int64 sum, the min and the max over the E envs.  Synchronises.        */
int32_t phx_reduce_field(phx_env* env, int32_t field, int32_t index, int32_t width, int32_t col,
                         int64_t* host_sum, int32_t* host_min, int32_t* host_max);

/* Resolver.tracked_messages (phantom/resolvers.py:41-60) of the LAST phx_step, for envs
 * [env_begin, env_end): host_counts int32 [n] messages recorded per env, host_msgs
 * int32 [n, trace_capacity, PHX_TRACE_WORDS] rows (sender_slot | recv_slot << 8 |
 * type << 16, payload0, payload1, round).  Needs PHX_FLAG_TRACK_MESSAGES.  Synchronises.
 * (After a T-step rollout this is the trace of its LAST step; see phx_get_trace_step.) */
int32_t phx_get_trace(phx_env* env, int32_t env_begin, int32_t env_end, int32_t* host_counts,
                      int32_t* host_msgs);

/* The same for step `step` (0-based) of the LAST tracked launch: a T-step phx_rollout of a handle
 * with PHX_FLAG_TRACK_MESSAGES records every step's messages (one slab of trace_capacity rows
 * per (step, env) on the device), which is what the reference's rollout utility collects into
 * Step.messages (phantom/utils/rollout.py:35-47, phantom/utils/rllib/rollout.py record_messages).
 * phx_trace_steps = T of that launch; phx_get_trace = its last step. */
int32_t phx_get_trace_step(phx_env* env, int32_t step, int32_t env_begin, int32_t env_end,
                           int32_t* host_counts, int32_t* host_msgs);
int32_t phx_trace_steps(const phx_env* env);

/* Run-time specialisation of the step kernel to ONE handle's env class.  The generic engines
 * interpret the lowered env class (agent counts, kinds, adjacency, stage tables) at run time;
 * phx_jit_source writes the text of a CUDA translation unit in which those are compile-time
 * constants (needs `buf_bytes` >= the text + NUL; *needed always receives the size; the unit
 * #includes files of phantom_b200/csrc, compile it with that directory on the include path:
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -I <csrc> unit.cu),
 * and phx_load_specialised loads the resulting cubin; from then on phx_step / phx_rollout of
 * this handle launch the specialised kernel (same results, bit for bit).  Supported by the
 * thread-per-env engine (env classes of at most 8 agents); PHX_ERR_UNSUPPORTED otherwise.  The
 * library contains no compiler: the host binding runs nvcc and caches cubins.            */
int32_t phx_jit_source(phx_env* env, char* buf, uint64_t buf_bytes, uint64_t* needed);
int32_t phx_load_specialised(phx_env* env, const char* cubin_path);

/* Collect device faults.  n_bad = envs whose error word is set, first_env / code = the
 * lowest such env and its phx_fault.  clear != 0 zeroes the words.  Synchronises.       */
int32_t phx_poll_errors(phx_env* env, int32_t* n_bad, int32_t* first_env, int32_t* code,
                        int32_t clear);

/* Self-test hook: evaluates the kernels' float32(n / den) routine (observation encoding,
 * supply_chain.py:124-134) for n = lo .. lo+count-1 on `device` into host_out float[count].
 * Lets tests compare the device arithmetic exhaustively against numpy's float32(n / den). */
int32_t phx_selftest_ratio(int32_t device, int32_t den, int32_t lo, int32_t count,
                           float* host_out);

/* Self-test hooks of the host side of phx_rollout_host (no GPU needed).  Families whose result
 * rows are small integers send them across PCIe in a compact wire format and expand them into
 * the caller's float32 planes on `threads` host threads (supply chain: one 32-bit word per
 * env-step, phantom_b200/csrc/phx_sc_wire.h; the planes replace what the reference computes in
 * encode_observation / compute_reward, supply_chain.py:124-147).  phx_selftest_wire_expand runs
 * that expansion on `n` words into obs float[n,3], reward float[n], all_done uint8[n,2];
 * phx_selftest_wire_pack builds one word. */
/* phx_jit_source for an env class WITHOUT a device: the text of the specialised translation unit
 * that a handle of `num_envs` envs of `spec` would get (build-time check that generated units
 * compile; see phx_jit_source). */
int32_t phx_selftest_jit_source(const phx_spec* spec, int32_t num_envs, uint64_t seed, char* buf,
                                uint64_t buf_bytes, uint64_t* needed);
int32_t phx_selftest_wire_expand(int32_t max_stock, int32_t cap, const uint32_t* wire, uint64_t n,
                                 int32_t threads, float* obs, float* reward, uint8_t* all_done);
uint32_t phx_selftest_wire_pack(int32_t stock, int32_t sales, int32_t missed, int32_t truncated,
                                int32_t was_reset);

#ifdef __cplusplus
}
#endif
#endif /* PHX_H_ */
