// snp_reset.cu -- SocialNavGym.reset for a whole batch on the device: every environment replays the reference's
// scenario generator (social_gym/social_nav_sim.py:200-431, chosen and seeded as social_gym/social_nav_gym.py:135-167 does) on its
// own copy of NumPy's MT19937 stream -- see snp_reset_core.h.  A WARP per environment: the rejection sampler is a sequential,
// data-dependent loop, so the 32 lanes SPECULATE on it -- lane l evaluates the l-th next attempt (the generator's block of 624
// words is random access), the first accepted one wins and the stream advances exactly as far as a sequential run would have -- and
// they share the 624-word twist of the generator (batches of 32 words).  Generator state and the placed humans live in the warp's slice of shared memory; results are
// written straight into the crowd's structure-of-arrays buffers.  Reset is not the hot path, but asynchronous episode ends make
// masked restarts frequent in an RL loop: 4096 envs x 25 humans restart in a fraction of a step's time, without leaving the GPU.
#include "snp_kernels.cuh"
#include "snp_reset_core.h"

namespace snp {
namespace {

template <typename T> struct ResetArgs {
    int E, N, G;
    long long EN;
    T *dyn, *stat, *goals, *robot;
    int *goal_idx, *goal_cnt;
    const uint32_t *seeds;
    uint32_t seed0;
    const uint8_t *mask;
    ResetParams p;
    double mass, robot_mass, robot_vd;
    double *time_now;
    int *flags;
    int32_t *scenario_out, *draws_out;
};

constexpr int kResetWarps = 4;

struct WarpGroup {
    SNP_HD int lane() const { return threadIdx.x & 31; }
    SNP_HD int size() const { return 32; }
    __device__ __forceinline__ int first(bool v) const { return __ffs(__ballot_sync(0xffffffffu, v)) - 1; }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

template <typename T> __global__ void __launch_bounds__(kResetWarps * 32) k_reset(const ResetArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = 624 * sizeof(uint32_t) + (size_t)4 * N * sizeof(double);
    uint32_t *mt = reinterpret_cast<uint32_t *>(smem_raw + warp * per_warp);     // [624]
    double *scr = reinterpret_cast<double *>(mt + 624);                          // [4][N]
    const long long env = (long long)blockIdx.x * kResetWarps + warp;
    if (env >= a.E) return;                       // whole warps leave together: the group operations below stay full-warp
    if (a.mask && !a.mask[env]) return;
    Mt19937<WarpGroup> rng{mt, 624, 0, WarpGroup{}};
    ResetScratch w{scr, scr + N, scr + 2 * N, scr + 3 * N};
    const long long EN = a.EN, base = env * N;
    const T mass = (T)a.mass;
    auto emit = [&](int i, const ResetHuman &h) {
        if (lane != 0) return;
        const long long k = base + i;
        a.dyn[SNP_DYN_PX * EN + k] = (T)h.x; a.dyn[SNP_DYN_PY * EN + k] = (T)h.y; a.dyn[SNP_DYN_TH * EN + k] = (T)h.yaw;
        a.dyn[SNP_DYN_VX * EN + k] = T(0); a.dyn[SNP_DYN_VY * EN + k] = T(0); a.dyn[SNP_DYN_BVX * EN + k] = T(0);
        a.dyn[SNP_DYN_BVY * EN + k] = T(0); a.dyn[SNP_DYN_OM * EN + k] = T(0); a.dyn[SNP_DYN_DFX * EN + k] = T(0);
        a.dyn[SNP_DYN_DFY * EN + k] = T(0);
        a.stat[SNP_STAT_R * EN + k] = (T)h.radius; a.stat[SNP_STAT_M * EN + k] = mass; a.stat[SNP_STAT_VD * EN + k] = (T)h.vd;
        a.goals[(size_t)0 * EN + k] = (T)h.g0x; a.goals[(size_t)1 * EN + k] = (T)h.g0y;
        if (a.G > 1) { a.goals[(size_t)2 * EN + k] = (T)h.g1x; a.goals[(size_t)3 * EN + k] = (T)h.g1y; }
        a.goal_idx[k] = 0;
        a.goal_cnt[k] = h.goal_count < a.G ? h.goal_count : a.G;
    };
    const uint32_t seed = a.seeds ? a.seeds[env] : a.seed0 + (uint32_t)env;
    const int scen = reset_generate(a.p, seed, rng, w, emit);
    if (lane != 0) return;
    if (a.robot) {  // sim:237 / :314: the robot of the scenario, at rest (social_nav_gym.py:213 robot.set(..., vx = 0, vy = 0))
        const long long E = a.E;
        T *r = a.robot + env;
        const bool pt = scen == SNP_SCEN_PARALLEL_TRAFFIC;
        const double half = a.p.traffic_length / 2, R = a.p.circle_radius;
        const double px = pt ? -half + 1 : 0.0, py = pt ? 0.0 : -R, gx = pt ? half - 1 : 0.0, gy = pt ? 0.0 : R;
        r[SNP_ROBOT_PX * E] = (T)px; r[SNP_ROBOT_PY * E] = (T)py; r[SNP_ROBOT_VX * E] = T(0); r[SNP_ROBOT_VY * E] = T(0);
        r[SNP_ROBOT_R * E] = (T)a.p.robot_radius; r[SNP_ROBOT_GX * E] = (T)gx; r[SNP_ROBOT_GY * E] = (T)gy;
        r[SNP_ROBOT_TH * E] = pt ? T(0) : (T)(3.141592653589793 / 2);
        r[SNP_ROBOT_BVX * E] = T(0); r[SNP_ROBOT_BVY * E] = T(0); r[SNP_ROBOT_OM * E] = T(0);
        r[SNP_ROBOT_M * E] = (T)a.robot_mass; r[SNP_ROBOT_VD * E] = (T)a.robot_vd; r[SNP_ROBOT_DFX * E] = T(0); r[SNP_ROBOT_DFY * E] = T(0);
        r[SNP_ROBOT_GX2 * E] = (T)px; r[SNP_ROBOT_GY2 * E] = (T)py; r[SNP_ROBOT_GCNT * E] = T(2);
    }
    if (a.time_now) a.time_now[env] = 0.0;  // social_nav_gym.py:129 global_time = 0
    if (a.flags) a.flags[env] = 0;
    if (a.scenario_out) a.scenario_out[env] = scen;
    if (a.draws_out) a.draws_out[env] = (int32_t)rng.draws;
}

template <typename T> int launch_reset(const snp_crowd *c, const snp_reset_args *g, cudaStream_t st) {
    ResetArgs<T> a;
    a.E = c->E; a.N = c->N; a.G = c->G; a.EN = (long long)c->E * c->N;
    a.dyn = (T *)c->dyn; a.stat = (T *)c->stat; a.goals = (T *)c->goals; a.robot = (T *)c->robot;
    a.goal_idx = c->goal_idx; a.goal_cnt = (int *)c->goal_cnt;
    a.seeds = g->seeds; a.seed0 = g->seed0; a.mask = g->mask;
    a.p.scenario = g->scenario; a.p.N = c->N; a.p.randomize_attributes = g->randomize_attributes;
    a.p.circle_radius = g->circle_radius; a.p.robot_radius = g->robot_radius;
    a.p.traffic_length = g->traffic_length; a.p.traffic_height = g->traffic_height;
    a.mass = g->human_mass; a.robot_mass = g->robot_mass; a.robot_vd = g->robot_desired_speed;
    a.time_now = g->time_now; a.flags = g->flags; a.scenario_out = g->scenario_out; a.draws_out = g->draws_out;
    const size_t smem = kResetWarps * (624 * sizeof(uint32_t) + (size_t)4 * c->N * sizeof(double));
    if (smem > 200 * 1024) { set_error("snp_reset: %d humans per env do not fit the generator's shared-memory scratch", c->N); return SNP_ERR_UNSUPPORTED; }
    auto kern = k_reset<T>;
    if (smem > 48 * 1024) SNP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)((c->E + kResetWarps - 1) / kResetWarps), kResetWarps * 32, smem, st>>>(a);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // namespace
}  // namespace snp

using namespace snp;

extern "C" int snp_reset(const snp_crowd *c, const snp_reset_args *g, void *stream) {
    if (!c || !g) { set_error("snp_reset: null descriptor"); return SNP_ERR_INVALID; }
    if (c->E <= 0 || c->N <= 0 || c->G <= 0) { set_error("snp_reset: E, N and G must be positive"); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->goals || !c->goal_idx || !c->goal_cnt) { set_error("snp_reset: crowd arrays missing"); return SNP_ERR_INVALID; }
    if (g->scenario < 0 || g->scenario > SNP_SCEN_HYBRID) { set_error("snp_reset: unknown scenario %d", g->scenario); return SNP_ERR_INVALID; }
    if ((g->scenario == SNP_SCEN_CCSO || g->scenario == SNP_SCEN_CCSO_SYNTHETIC) && c->N < 3) { set_error("snp_reset: the static-obstacle scenarios need at least 3 humans"); return SNP_ERR_INVALID; }
    if (g->scenario != SNP_SCEN_PARALLEL_TRAFFIC && c->G < 2) { set_error("snp_reset: the circular scenarios need two goal slots per human"); return SNP_ERR_INVALID; }
    {   // generate_parallel_traffic_scenario raises when the humans cannot fit (social_nav_sim.py:327-329); worst case radius 0.5
        const double r = g->randomize_attributes ? 0.5 : 0.3;
        if ((g->scenario == SNP_SCEN_PARALLEL_TRAFFIC || g->scenario == SNP_SCEN_HYBRID) &&
            c->N * 3.141592653589793 * r * r > g->traffic_length * g->traffic_height * 0.4) {
            set_error("Number of humans specified is too big for desided traffic height and length");
            return SNP_ERR_INVALID;
        }
    }
    if (c->dtype == SNP_F64) return launch_reset<double>(c, g, (cudaStream_t)stream);
    if (c->dtype == SNP_F32) return launch_reset<float>(c, g, (cudaStream_t)stream);
    set_error("dtype %d is neither SNP_F32 nor SNP_F64", c->dtype);
    return SNP_ERR_INVALID;
}
