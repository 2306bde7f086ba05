// snp_kernels.cuh -- argument blocks shared by the kernels and the C-ABI glue.
#pragma once
#include <cstdint>
#include "snp_physics.cuh"
#include "../../include/snp_b200.h"

namespace snp {

template <typename T> struct alignas(16) Ent { T x, y, vx, vy; };  // float4 / double4-sized entity record staged in shared memory

// Entities of one env group in shared memory as two arrays of (x,y) and (vx,vy) pairs.  With one lane per entity (the halved
// pair loop reads ents[(i+k) mod N]) consecutive lanes then touch consecutive 8/16-byte words -> no bank conflicts, where a
// 32-byte double4 record per lane gave a 2-way conflict on every LDS.128 (profiles/r01: 11.3 M conflicts per launch).
template <typename T> struct alignas(2 * sizeof(T)) Vec2 { T a, b; };
template <typename T> struct EntView {
    Vec2<T> *pos, *vel;
    __device__ __forceinline__ Ent<T> get(int j) const { const Vec2<T> p = pos[j], v = vel[j]; return Ent<T>{p.a, p.b, v.a, v.b}; }
    __device__ __forceinline__ void put(int j, T x, T y, T vx, T vy) const { pos[j] = Vec2<T>{x, y}; vel[j] = Vec2<T>{vx, vy}; }
};

template <typename T> struct KArgs {
    int E, N, G;
    long long EN;
    T *dyn;
    const T *stat;
    const T *goals;
    int *goal_idx;
    const int *goal_cnt;
    const T *agent_params;
    Params<T> P;
    T *robot;
    const T *walls;
    int W, S, walls_per_env;
    int consider_robot, symmetric, numba, n_substeps, robot_mode;
    T dt;
    double dt_d;
    const T *action;
    int pre_checks, post_checks, track_touch;
    double consts[6];
    double *time_now;
    int *flags;
    double *checks;
    int epw;  // envs per warp (warp-packed kernel)
    int robot_type;           // robot_mode 2: the robot's model
    Params<T> RP;             // robot_mode 2: the robot's parameters
    int respawn;              // parallel-traffic respawn after every sub-step (mmm:407-422)
    double respawn_bounds[2];
    int gpb;  // env groups per block (block-packed kernel)
    int mapping;  // 0 auto, 1 warp-packed, 2 block-packed
    int full_pair_loop;  // 1: always evaluate ordered pairs in j-ascending order (the reference's accumulation order)
    const int *respawn_envs;  // optional per-env respawn switch
    int *goal_idx_out;  // peek: goal index after the update (optional)
    T *dyn_out;  // peek: updated pose / velocities are written here instead of in place (goal index untouched)
    int robot_every, robot_phase;  // robot_mode 2: SocialNavSim.update schedule (0 = imitation-learning order), see snp_step_opts
    T robot_dt;  // consts[5] in the crowd's dtype
    // snp_gym_step_host with pinned (device-mapped) host buffers: the store phase writes the observation (px, py, vx, vy planes),
    // flags and checks STRAIGHT into host memory -- posted PCIe writes that overlap the rest of the launch; no D2H copy follows
    T *obs_out = nullptr;
    int *flags2 = nullptr;
    double *checks2 = nullptr;
};

// Host-side launchers implemented per translation unit.
template <typename T> int launch_step_small(const KArgs<T> &a, int type, cudaStream_t st);
template <typename T> int launch_checks(const KArgs<T> &a, cudaStream_t st);

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int device_sm_count();

#define SNP_CUDA_OK(expr)                                                                 \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            snp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SNP_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

}  // namespace snp
