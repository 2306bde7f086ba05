/*
 * snp_b200.h -- C ABI of libsnp_b200.so: the B200 (sm_100a) batched crowd-stepping engine that drops in for the
 * per-step human motion update of Social-Navigation-PyEnvs.
 *
 * Everything here is plain C: pointers, ints, doubles.  No torch / C++ types cross this boundary.  Device-pointer
 * entry points are stream-ordered, allocate nothing and never synchronise; host-pointer entry points (suffix _host)
 * copy in, run the same kernels, copy out and synchronise the stream they use.
 *
 * Every function returns 0 on success or a negative snp_status; snp_last_error() gives the message
 * (the Python adapter turns it into ValueError / RuntimeError, matching the reference's exceptions,
 * social_gym/src/forces_parallel.py:211).
 *
 * Reference interfaces replaced (paths relative to the reference's social_gym/):
 *   snp_step / snp_update_humans_parallel_host   <- src/forces_parallel.py:184-284 update_humans_parallel  (sole call
 *                                                   site src/motion_model_manager.py:360), and the serial path it
 *                                                   shadows: motion_model_manager.py:354-373,424-459 + src/forces.py
 *   fused sub-step loop + robot motion            <- social_nav_gym.py:240-245 (20 x robot.step + update_humans),
 *                                                   src/robot_agent.py:126-136
 *   snp_gym_step_host                             <- social_nav_gym.py:227-250 SocialNavGym.step with host action / observation buffers
 *   snp_checks (+ fused into snp_step)            <- social_nav_sim.py:949-984 collision_detection_and_reaching_goal,
 *                                                   :986-1029 compute_reward_and_infos, :702-703;
 *                                                   social_nav_gym.py:107-118 check_actual_collisions_and_goal
 *   snp_laser / snp_laser_host                    <- src/sensors.py:53-69 LaserSensor.get_laser_measurements
 *                                                   (:24-33 circle, :35-51 segment), src/robot_agent.py:77-82
 *   respawn option of snp_step                    <- motion_model_manager.py:407-422 (parallel-traffic post_update)
 *   robot_mode 2 of snp_step                      <- motion_model_manager.py:593-653 update_robot / compute_robot_forces,
 *                                                   social_nav_gym.py:252-274 imitation_learning_step
 *   snp_lookahead (+ dyn_out peek of snp_step)    <- crowd_nav/policy/cadrl.py:42-83 compute_rotated_states_and_reward, :13-40;
 *                                                   motion_model_manager.py:691-709 get_next_human_observable_states
 *   snp_robot_push_out                            <- src/robot_agent.py:35-48 RobotAgent.check_collisions (social_nav_sim.py:509)
 *   snp_reset                                     <- social_nav_gym.py:120-225 reset, social_nav_sim.py:200-431 scenario generators
 *   snp_pack_states / snp_unpack_states           <- src/agent.py:256-266 get_safe_state / set_state row layout
 *   snp_large_step                                <- same update for one very large crowd (tiled all-pairs)
 */
#ifndef SNP_B200_H
#define SNP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNP_ABI_VERSION 3

typedef enum snp_status {
    SNP_OK = 0,
    SNP_ERR_INVALID = -1, /* bad argument (type outside 0..8, sizes, null pointers) -> ValueError */
    SNP_ERR_CUDA = -2,    /* CUDA runtime error */
    SNP_ERR_UNSUPPORTED = -3
} snp_status;

enum { SNP_F32 = 0, SNP_F64 = 1 };
enum { SNP_OPT_FULL_PAIR_LOOP = 1, SNP_OPT_NO_CULLING = 2, SNP_OPT_MAP_WARP = 4, SNP_OPT_MAP_BLOCK = 8,
       SNP_OPT_STAGED_COPIES = 16, /* snp_gym_step_host: D2H copies after the launch even for pinned result buffers */
       SNP_OPT_LARGE_GRID = 32     /* snp_large_*: culled steps on the static (i-block, chunk) grid instead of the work list (tests / tuning) */ };

/* Field order of the structure-of-arrays buffers (each field is a contiguous run of E*N elements). */
enum { SNP_DYN_PX = 0, SNP_DYN_PY, SNP_DYN_VX, SNP_DYN_VY, SNP_DYN_TH, SNP_DYN_BVX, SNP_DYN_BVY, SNP_DYN_OM,
       SNP_DYN_DFX, SNP_DYN_DFY, SNP_DYN_FIELDS };                        /* dfx,dfy: carried desired force (forces.py:12-15) */
enum { SNP_STAT_R = 0, SNP_STAT_M, SNP_STAT_VD, SNP_STAT_SAFETY, SNP_STAT_FIELDS };
enum { SNP_ROBOT_PX = 0, SNP_ROBOT_PY, SNP_ROBOT_VX, SNP_ROBOT_VY, SNP_ROBOT_R, SNP_ROBOT_SAFETY, SNP_ROBOT_GX,
       SNP_ROBOT_GY, SNP_ROBOT_TH,
       /* only used when the robot is driven by a motion model (robot_mode == 2): */
       SNP_ROBOT_BVX, SNP_ROBOT_BVY, SNP_ROBOT_OM, SNP_ROBOT_M, SNP_ROBOT_VD, SNP_ROBOT_DFX, SNP_ROBOT_DFY,
       SNP_ROBOT_GX2, SNP_ROBOT_GY2,   /* the other goal of the robot's (at most two-entry) goal list */
       SNP_ROBOT_GCNT,                 /* 1 or 2 goals */
       SNP_ROBOT_SPARE, SNP_ROBOT_FIELDS };

/* Bits of the per-env flags word written by snp_step / snp_checks. */
enum {
    SNP_FLAG_COLLISION = 1 << 0,        /* swept test over one robot step (social_nav_sim.py:949-979) */
    SNP_FLAG_REACHING_GOAL = 1 << 1,    /* social_nav_sim.py:980-983 */
    SNP_FLAG_TERMINATED = 1 << 2,       /* social_nav_sim.py:986-1029 */
    SNP_FLAG_TRUNCATED = 1 << 3,
    SNP_FLAG_INFO_SHIFT = 4,            /* 3 bits: 0 Nothing 1 Timeout 2 Collision 3 ReachGoal 4 Danger */
    SNP_FLAG_ACTUAL_COLLISION = 1 << 7, /* social_nav_gym.py:107-118 on the post-step state */
    SNP_FLAG_ACTUAL_GOAL = 1 << 8,
    SNP_FLAG_TOUCHED = 1 << 9           /* |p_h - p_r| < r_h + r_r after any sub-step (social_nav_sim.py:702-703) */
};

/* A batch of E independent environments with N humans each, resident in device memory.
 * `dtype` selects float or double for every `void*` array below. */
typedef struct snp_crowd {
    int32_t E, N, G;          /* envs, humans per env, goal slots per human */
    int32_t dtype;            /* SNP_F32 / SNP_F64 */
    void *dyn;                /* [SNP_DYN_FIELDS][E*N]   updated in place */
    const void *stat;         /* [SNP_STAT_FIELDS][E*N] */
    const void *goals;        /* [G][2][E*N]  goal lists; slots >= goal_cnt are ignored */
    int32_t *goal_idx;        /* [E*N] index of the current goal (the reference rotates the list instead,
                                 motion_model_manager.py:66-70); updated in place */
    const int32_t *goal_cnt;  /* [E*N] number of valid goals (>= 1) */
    const void *agent_params; /* optional [20][E*N] per-agent parameter rows (agent.py:269); NULL -> `params` for all */
    double params[20];        /* uniform parameter row used when agent_params == NULL */
    void *robot;              /* optional [SNP_ROBOT_FIELDS][E]; required when consider_robot or checks are on */
    const void *walls;        /* optional [walls_per_env ? E : 1][W*S][4] = ax,ay,bx,by per segment slot, NaN-padded
                                 (motion_model_manager.py:268-276), endpoints ordered as obstacle.py:31-32 */
    int32_t W, S;             /* wall polygons, segment slots per polygon */
    int32_t walls_per_env;
} snp_crowd;

typedef struct snp_step_opts {
    int32_t type;             /* 0..8 index into SFMS (motion_model_manager.py:15-17) */
    int32_t consider_robot;   /* robot exerts force on humans (motion_model_manager.py:35) */
    int32_t symmetric;        /* all_equal_humans: pair (i,j), i<j evaluated with i as agent1 and applied +/- (forces.py:130-151);
                                 0 -> per-agent path (forces.py:153-218) */
    int32_t numba_compat;     /* 0: serial Python/NumPy semantics (oracle of record); 1: forces_parallel.py semantics
                                 ('<=' goal test, zeroed desired force, Guo wall force / W, first-wins closest segment) */
    int32_t n_substeps;       /* fused update_humans calls per launch (20 in SocialNavGym.step) */
    int32_t robot_mode;       /* 0: robot row fixed during the launch; 1: holonomic action: before every sub-step
                                 p += a*dt, v = a (robot_agent.py:126-131); 2: the robot is moved by its own SFM / HSFM model
                                 before every human update (motion_model_manager.py:593-653 update_robot, as in
                                 SocialNavGym.imitation_learning_step, social_nav_gym.py:260-265); 3: unicycle action (v, r):
                                 before every sub-step p += v (cos, sin)(yaw + r) dt, yaw = (yaw + r) % 2 pi, velocity along the
                                 new yaw (robot_agent.py:116-136); the swept check uses v (cos, sin)(yaw + r) (social_nav_sim.py:973,
                                 with the robot's yaw where the reference reads its never-assigned `theta`) */
    double dt;
    const void *action;       /* [2][E] (dtype of the crowd) when robot_mode == 1 / 3 or pre_checks: (vx, vy), or (v, r) in mode 3 */
    int32_t pre_checks;       /* swept collision / goal / reward on the PRE-step state (social_nav_gym.py:232-234) */
    int32_t post_checks;      /* 1: actual collision / goal on the POST-step state (social_nav_gym.py:107-118); 2: also reward, terminated,
                                 truncated and info code from them at the end time (imitation_learning_step, social_nav_gym.py:269-271) */
    int32_t track_touch;      /* OR of the run_k_steps collision test after every sub-step */
    int32_t reserved;         /* bit 0 (SNP_OPT_FULL_PAIR_LOOP): evaluate every ordered pair in j-ascending order (the reference's
                                 accumulation order) instead of each unordered pair once per warp;
                                 bits 2-3 (SNP_OPT_MAP_WARP / SNP_OPT_MAP_BLOCK): force the warp-packed / block-packed thread mapping */
    double consts[6];         /* time_limit, collision_penalty, success_reward, discomfort_dist,
                                 discomfort_penalty_factor, robot_time_step */
    double *time_now;         /* optional [E] global_time; read by the reward, advanced by dt per sub-step */
    int32_t *flags;           /* [E] out, required when any check is on */
    double *checks;           /* [E][4] out: dmin (swept), reward, dmin (actual), unused */
    double respawn_bounds[2]; /* (traffic_length / 2, traffic_height / 2)  (social_nav_sim.py:360) */
    int32_t respawn;          /* parallel-traffic respawn after every sub-step (motion_model_manager.py:407-422): humans within 3 m of
                                 their goal restart at the right end; rewrites goals[0], goal_cnt (N <= 32 only) */
    int32_t robot_type;       /* robot_mode 2: the robot's model, 0..8 (may differ from the humans') */
    double robot_params[20];  /* robot_mode 2: the robot's parameter row (agent.py:269 for its model) */
    void *dyn_out;            /* optional [SNP_DYN_FIELDS][E*N]: PEEK (motion_model_manager.py:691-709 get_next_human_observable_states):
                                 the updated px..omega fields are written here instead of in place and the goal index is not
                                 advanced; the carried desired force is still updated in place (the reference does not restore it) */
    const int32_t *respawn_envs; /* optional [E], with respawn: only envs with a non-zero entry respawn (hybrid scenario: the parallel-traffic
                                 envs of a mixed batch; snp_reset's scenario_out works as is) */
    int32_t *goal_idx_out;    /* optional [E*N], with dyn_out: the goal index the update arrived at (the peek reports the goal after it) */
    int32_t robot_every;      /* robot_mode 2 only.  0: SocialNavGym.imitation_learning_step order (update_robot, then humans that see the
                                 moved robot).  >= 1: SocialNavSim.update / control_robot (social_nav_sim.py:476-529): the humans see the
                                 robot's state from BEFORE its update (:484-491); 1 = equal sampling times, update_robot(dt) every
                                 sub-step (:521); k > 1 = the pose advances every sub-step with the last velocity, yaw unwrapped
                                 (update_robot_pose, motion_model_manager.py:655-657) and every k-th sub-step the velocities are refreshed
                                 by update_robot(..., consts[5] = ROBOT_SAMPLING_TIME, just_velocities=True) (:523-524, mmm:72-85) */
    int32_t robot_phase;      /* robot_every > 1: index of this launch's first sub-step in that schedule (n_updates of the simulator) */
} snp_step_opts;

typedef struct snp_laser_args {
    int32_t E, N;             /* envs, circles (humans) per env */
    int32_t dtype;
    int32_t samples;
    const void *px, *py, *radius; /* [E*N] each (SoA fields of a crowd work as is) */
    const void *walls;        /* as snp_crowd.walls */
    int32_t W, S, walls_per_env;
    int32_t reserved;
    const void *pose;         /* [3][E] x, y, yaw of the sensor */
    double range, max_distance, robot_radius; /* robot_radius is subtracted from the ranges (robot_agent.py:81); 0 for raw */
    void *ranges;             /* [E][samples] out */
    int32_t *hits;            /* [E][samples] out (optional): human index, N + segment ordinal, or -1 */
    /* ABI 3 (appended): LaserSensor.add_uncertainty (sensors.py:71-74) on the device -- every range becomes
     * clip(N(range, uncertainty), 0, max_distance) before the robot radius is subtracted.  The reference draws from the caller's global
     * np.random stream, one normal per ray in ray order; here ray (env, k) of scan `noise_scan` takes its normal from the counter-based
     * Philox4x32-10 stream keyed by `noise_seed` (counter = k, env, scan): same distribution, reproducible, independent of the launch
     * geometry.  uncertainty <= 0 (or NaN): no noise, as with uncertainty=None. */
    double uncertainty;
    uint64_t noise_seed;
    uint64_t noise_scan;
} snp_laser_args;

/* One-step lookahead of the value-network policies (crowd_nav/policy/cadrl.py:42-83 compute_rotated_states_and_reward,
 * called from CADRL.predict :235-262) for every env of a crowd and every action of the action space. */
typedef struct snp_lookahead_args {
    int32_t type;             /* the humans' model 0..8 (headed models take yaw / omega from `next`, the others from the crowd) */
    int32_t A;                /* number of actions */
    int32_t theta_and_omega_visible; /* 0: 13 values per (action, human); 1: 15 (cadrl.py:14-21) */
    int32_t reserved;         /* bit 0: write the rows with per-thread vector stores instead of bulk asynchronous copies (A/B measurement) */
    const void *next;         /* [SNP_DYN_FIELDS][E*N] the humans one policy time step ahead: output of the peek (snp_step with dyn_out) */
    const double *actions;    /* [A][2] holonomic velocities (vx, vy), shared by all envs (cadrl.py build_action_space) */
    double dt;                /* the policy's time step (0.25) */
    void *rotated;            /* out [E][A][N][13|15] in the crowd's dtype: dg, v_pref, theta, radius, vx, vy, px1, py1, vx1, vy1, radius1, da,
                                 radius_sum (, theta1, omega1) */
    double *rewards;          /* out [E][A]: -0.25 collision, 1 goal, (dmin - 0.2) * 0.5 * dt discomfort, 0 (cadrl.py:68-72), bit-exact */
} snp_lookahead_args;

/* SocialNavGym.reset for every (selected) environment of a crowd, on the device (social_gym/social_nav_gym.py:120-225): env e replays
 * the reference's scenario generator (social_gym/social_nav_sim.py:200-299 circular crossing, :301-362 parallel traffic, :364-431
 * circular crossing with static obstacles; insert_robot = True, randomize_human_positions = True) on NumPy's own MT19937 stream seeded
 * with seeds[e] (the gym seeds np.random with offset[phase] + case, social_nav_gym.py:135-137): same draws, same accept / reject
 * decisions, positions equal up to the last ulp of cos / sin. */
enum { SNP_RESET_CIRCULAR_CROSSING = 0, SNP_RESET_PARALLEL_TRAFFIC = 1, SNP_RESET_CCSO = 2,
       SNP_RESET_CCSO_SYNTHETIC = 3, /* 3 static humans + the circular-crossing sampler: the 25-human crowd of SURVEY.md 8(d) config 3 */
       SNP_RESET_HYBRID = 4          /* np.random.choice between the first two, then re-seed (social_nav_gym.py:155-157) */ };
typedef struct snp_reset_args {
    int32_t scenario;
    int32_t randomize_attributes; /* desired speed ~ U(0.5,1.5), radius ~ U(0.3,0.5) per human, drawn first (social_nav_sim.py:217-220) */
    const uint32_t *seeds;        /* optional [E] device: np.random.seed argument of each env; NULL -> seed0 + env index */
    uint32_t seed0;
    uint32_t reserved;
    const uint8_t *mask;          /* optional [E] device: only envs with a non-zero entry are reset (restart of finished episodes) */
    double circle_radius, robot_radius, traffic_length, traffic_height;
    double human_mass, robot_mass, robot_desired_speed; /* 75 (social_nav_sim.py:141), 80 (robot_agent.py:16), 1 */
    double *time_now;             /* optional [E]: set to 0 (social_nav_gym.py:129) */
    int32_t *flags;               /* optional [E]: cleared */
    int32_t *scenario_out;        /* optional [E]: the scenario generated (the hybrid scenario's coin) */
    int32_t *draws_out;           /* optional [E]: number of uniforms the generator consumed */
} snp_reset_args;

int snp_abi_version(void);
const char *snp_last_error(void);
/* Device properties the host side sizes grids with: sm_count, cc_major, cc_minor, l2_bytes. */
int snp_device_info(int32_t *out4);

/* ---- device-pointer API ---- */
int snp_step(const snp_crowd *crowd, const snp_step_opts *opts, void *cuda_stream);
int snp_checks(const snp_crowd *crowd, const snp_step_opts *opts, void *cuda_stream);
int snp_laser(const snp_laser_args *args, void *cuda_stream);
/* RobotAgent.check_collisions (src/robot_agent.py:35-48): crowd->robot's position is pushed out of the humans it overlaps (in index
 * order), then out of the wall polygons it overlaps; sequential per env, bit-identical to the reference for fp64 state. */
int snp_robot_push_out(const snp_crowd *crowd, void *cuda_stream);
/* Rewrites crowd->dyn, stat (radius, mass, desired speed; the safety space is kept), goals, goal_idx, goal_cnt and crowd->robot. */
int snp_reset(const snp_crowd *crowd, const snp_reset_args *args, void *cuda_stream);
/* Uses crowd->dyn (current px,py,vx,vy,theta,omega), crowd->stat (radius) and crowd->robot (px,py,r,gx,gy,vd). */
int snp_lookahead(const snp_crowd *crowd, const snp_lookahead_args *args, void *cuda_stream);
/* The policies' query_env = False branch (crowd_nav/policy/cadrl.py:92-105 propagate_humans_state_with_constant_velocity_model):
 * `next` [SNP_DYN_FIELDS][E*N] (crowd dtype) = the humans one step of length dt ahead under constant velocity (x + vx dt, y + vy dt,
 * theta + omega dt, velocities carried over) -- the `next` input of snp_lookahead when the env is not queried. */
int snp_constant_velocity(const snp_crowd *crowd, double dt, void *next, void *cuda_stream);
/* AoS <-> SoA: rows are the reference's 13-wide float64 state rows [E][rows][13] with the robot (if any) as row N. */
int snp_unpack_states(const snp_crowd *crowd, const double *rows_dev, int32_t rows_per_env, const double *safety_dev,
                      void *cuda_stream);
int snp_pack_states(const snp_crowd *crowd, double *rows_dev, int32_t rows_per_env, void *cuda_stream);
/* Goal lists: reference rows [E][N][G][2] NaN padded (motion_model_manager.py:262-267) -> crowd->goals / goal_idx (= 0) and
 * goal_cnt_dev (writable alias of crowd->goal_cnt).  snp_rotate_goal_rows applies the rotation the reference would have
 * performed in place (forces_parallel.py:229-232) to the caller's float64 rows, given the crowd's goal_idx. */
int snp_unpack_goals(const snp_crowd *crowd, const double *goal_rows_dev, int32_t *goal_cnt_dev, void *cuda_stream);
int snp_rotate_goal_rows(const snp_crowd *crowd, double *goal_rows_dev, void *cuda_stream);
/* One very large crowd (the crowd's E*N agents form ONE environment): one sub-step, shared-memory tiled all-pairs.
 * `others` = [5][M] x,y,vx,vy,r+safety (SoA) of ALL M entities exerting force -- the crowd itself (gathered over ranks
 * when it is sharded by agent) plus, optionally, the robot as last entry; `self_offset` = index of this crowd's agent 0
 * in that view.  Own state is updated in place; when `next_view` is non-NULL the updated x,y,vx,vy,r+safety of the own
 * agents are also written into it (same [5][M] layout, same offset), ready to be all-gathered for the next sub-step.
 * `scratch` = device workspace of at least snp_large_scratch_bytes(n_local, M, dtype) bytes (per-chunk partial sums, tile
 * boxes, the map and lists of the (agent block, entity chunk) units in reach, the work queues' counters).  opts->reserved bit 1
 * (SNP_OPT_NO_CULLING) disables the exact far-tile culling; SNP_OPT_LARGE_GRID runs a culled step on the static grid of units
 * instead of the list worked off by persistent CTAs (same results bit for bit, slower; tests / tuning). */
int64_t snp_large_scratch_bytes(int64_t n_local, int64_t M, int32_t dtype);
int snp_large_step(const snp_crowd *crowd, const snp_step_opts *opts, const void *others, int64_t M, int64_t self_offset,
                   void *next_view, void *scratch, int64_t scratch_bytes, void *cuda_stream);
/* Same sub-step with the all-gather FUSED into the producer kernel: `peer_next_views` is a host array of `n_peers` device
 * pointers -- the NEXT-view buffer of every rank of the node (this rank's own included), peer-mapped over NVLink (e.g. torch
 * symmetric memory) -- and the finish kernel stores each agent's new entry straight into all of them.  The caller only needs a
 * cross-rank barrier between sub-steps; no collective moves data. */
int snp_large_step_p2p(const snp_crowd *crowd, const snp_step_opts *opts, const void *others, int64_t M, int64_t self_offset,
                       const void *const *peer_next_views, int32_t n_peers, void *scratch, int64_t scratch_bytes, void *cuda_stream);
/* `n_substeps` of the agent-sharded step in ONE call (the loop SocialNavGym.step runs, social_nav_gym.py:240-245, for one crowd spread
 * over `world` GPUs of a node): per sub-step two launches (tiled pairs, finish) + one single-warp barrier kernel, all enqueued on
 * `cuda_stream` without returning to the host.  Buffers, all peer-mapped (e.g. torch symmetric memory), as host arrays of `world`
 * device pointers indexed by rank: `peer_views_a` / `peer_views_b` = the two view buffers, each [5][M] entities followed by the
 * view's [ceil(M/128)][5] tile-box table (the producer writes entries AND boxes into every rank's next buffer); `peer_flags` =
 * uint64[world] barrier slots per rank, zero-initialised once.  `first_is_b` selects the buffer that holds the current view;
 * `epoch_base` must grow by `n_substeps` from call to call.  `error_flag` (device int32) is set if a peer never reached a barrier. */
int snp_large_run_p2p(const snp_crowd *crowd, const snp_step_opts *opts, const void *const *peer_views_a, const void *const *peer_views_b,
                      int32_t first_is_b, int64_t M, int64_t self_offset, int32_t world, int32_t rank, const void *const *peer_flags,
                      uint64_t epoch_base, int32_t n_substeps, int32_t *error_flag, void *scratch, int64_t scratch_bytes, void *cuda_stream);
/* Writes this crowd's [5][.] entity view (x,y,vx,vy,r+safety; v = R(yaw) bv for headed models) into `view` (field stride
 * `stride` elements, starting at element `offset`). */
int snp_large_publish(const snp_crowd *crowd, int32_t type, void *view, int64_t stride, int64_t offset, void *cuda_stream);

/* ---- host-pointer API (the reference operator, batched over a leading env axis) ----
 * update_humans_parallel(type, agents_state, goals, obstacles, agents_params, dt, safety_space,
 *                        all_params_equal, last_is_robot) -> updated_state        forces_parallel.py:184
 * agents_state [E][rows][13] (rows = N + last_is_robot), goals [E][N][G][2] NaN-padded and ROTATED IN PLACE on goal
 * reach, agents_state[:,:,10:12] refreshed and (headed models) [:,:,3:5] overwritten as the reference does
 * (fp:229-234,256); obstacles [W][S][2][2] NaN-padded or NULL; agents_params [E][N][20]; safety_space [E][rows];
 * desired_force [E][N][2] in/out or NULL (zeros; only the serial semantics carry it); out_state [E][rows][13]. */
int snp_update_humans_parallel_host(int32_t type, int32_t E, int32_t N, int32_t G, double *agents_state, double *goals,
                                    const double *obstacles, int32_t W, int32_t S, const double *agents_params, double dt,
                                    const double *safety_space, int32_t all_params_equal, int32_t last_is_robot,
                                    int32_t numba_compat, int32_t dtype, int32_t n_substeps, double *desired_force,
                                    double *out_state);
/* SocialNavGym.step (social_gym/social_nav_gym.py:227-250) for a crowd that stays RESIDENT on the device, with host buffers for what
 * crosses the boundary every step: `action_host` [2][E] (crowd dtype; copied into opts->action), then snp_step, then the observation
 * `obs_host` [4][E*N] = px, py, vx, vy of every human (crowd dtype), `flags_host` [E] and `checks_host` [E][4] (dmin, reward, actual
 * dmin, -) are copied back and the stream is synchronised.  Pass pinned host memory for asynchronous copies; any of the host
 * pointers may be NULL. */
int snp_gym_step_host(const snp_crowd *crowd, const snp_step_opts *opts, const void *action_host, void *obs_host, int32_t *flags_host,
                      double *checks_host, void *cuda_stream);
/* LaserSensor.get_laser_measurements over E sensors: humans [E][N][3] = x,y,r; walls [W][S][2][2]; pose [E][3];
 * ranges [E][samples]; hits [E][samples] or NULL. */
int snp_laser_host(int32_t E, int32_t N, const double *humans, const double *walls, int32_t W, int32_t S, const double *pose,
                   double range, int32_t samples, double max_distance, double robot_radius, int32_t dtype, double *ranges,
                   int32_t *hits);

/* ---- measurement helpers (bench.py) ---- */
/* Runs a register-resident FMA / MUFU loop on every SM and returns achieved TFLOP/s (kind 0 fp32 FMA, 1 fp64 FMA)
 * or Gop/s (kind 2 MUFU.EX2).  Used as the measured denominator of the compute roofline. */
int snp_measure_pipe_peak(int32_t kind, double *out);
/* Test hook: y[i] = the kernels' table-based fp64 exp (csrc/snp_math.cuh exp_tbl) of x[i]; device pointers. */
int snp_debug_exp(const double *x_dev, double *y_dev, int32_t n, void *cuda_stream);
/* Test hook for the kernels' own fp64 elementary functions (csrc/snp_math.cuh); device pointers, n elements each.
 * kind 0: out[i] = exp2_scaled(x[i]) = 2^(x[i] / 2048);  1: out[i] = atan2_poly(y[i], x[i]);
 * 2: out[i] = sin, out[n + i] = cos of x[i] by sincos_bounded (|x| <= pi + 1);  3: out[i] = rsqrt_(x[i]);  4: out[i] = clamp01_(x[i]). */
int snp_debug_math(int32_t kind, const double *x_dev, const double *y_dev, double *out_dev, int32_t n, void *cuda_stream);
/* Launch statistics since the last reset: number of kernels this library launched. */
int64_t snp_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* SNP_B200_H */
