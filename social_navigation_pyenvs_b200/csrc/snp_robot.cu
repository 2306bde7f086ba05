// snp_robot.cu -- RobotAgent.check_collisions for every env (social_gym/src/robot_agent.py:35-48; called from
// SocialNavSim.control_robot, social_nav_sim.py:509): the robot is pushed out of every human it overlaps, in list order, then out of
// every wall polygon it overlaps (closest point of src/obstacle.py:53-66), each push seeing the result of the previous ones.
// The chain is sequential and order-dependent, so it is one thread per env; every operation is the reference's own (np.linalg.norm /
// np.dot in their OpenBLAS FMA forms, IEEE divide, no other contraction), which makes the result bit-identical for fp64 state.
#include "snp_kernels.cuh"

namespace snp {
namespace {

template <typename T> struct PushArgs {
    int E, N, W, S, walls_per_env;
    long long EN;
    const T *dyn, *stat, *walls;
    T *robot;
};

template <typename T> __global__ void k_robot_push_out(const PushArgs<T> a) {
    const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= a.E) return;
    const long long E = a.E, EN = a.EN, base = env * a.N;
    double rx = (double)a.robot[SNP_ROBOT_PX * E + env], ry = (double)a.robot[SNP_ROBOT_PY * E + env];
    const double rr = (double)a.robot[SNP_ROBOT_R * E + env];
    for (int j = 0; j < a.N; ++j) {  // robot_agent.py:36-41
        const double hx = (double)a.dyn[SNP_DYN_PX * EN + base + j], hy = (double)a.dyn[SNP_DYN_PY * EN + base + j];
        const double hr = (double)a.stat[SNP_STAT_R * EN + base + j];
        const double dx = __dsub_rn(rx, hx), dy = __dsub_rn(ry, hy);
        const double dist = xnorm_np(dx, dy), sum = __dadd_rn(hr, rr);
        if (dist < sum) {
            rx = __dadd_rn(hx, __dmul_rn(__ddiv_rn(dx, dist), sum));
            ry = __dadd_rn(hy, __dmul_rn(__ddiv_rn(dy, dist), sum));
        }
    }
    for (int w = 0; w < a.W; ++w) {  // robot_agent.py:42-47 with obstacle.py:53-66 ('<=': the last of equally close segments wins)
        const T *wall = a.walls + ((size_t)(a.walls_per_env ? env : 0) * a.W + w) * a.S * 4;
        double best = 10000.0, cx = 0.0, cy = 0.0;
        for (int s = 0; s < a.S; ++s) {
            const double ax = (double)wall[4 * s];
            if (ax != ax) continue;  // NaN padding
            const double ay = (double)wall[4 * s + 1];
            const double ex = __dsub_rn((double)wall[4 * s + 2], ax), ey = __dsub_rn((double)wall[4 * s + 3], ay);
            const double len = xnorm_np(ex, ey);
            const double t = __ddiv_rn(xdot_np(__dsub_rn(rx, ax), __dsub_rn(ry, ay), ex, ey), __dmul_rn(len, len));
            double ts = t > 0.0 ? t : 0.0;
            ts = 1.0 < ts ? 1.0 : ts;
            const double hx = __dadd_rn(ax, __dmul_rn(ts, ex)), hy = __dadd_rn(ay, __dmul_rn(ts, ey));
            const double d = xnorm_np(__dsub_rn(hx, rx), __dsub_rn(hy, ry));
            if (d <= best) { best = d; cx = hx; cy = hy; }
        }
        if (best < rr) {
            const double nn = xnorm_np(__dsub_rn(cx, rx), __dsub_rn(cy, ry));
            const double ux = __ddiv_rn(__dsub_rn(rx, cx), nn), uy = __ddiv_rn(__dsub_rn(ry, cy), nn);
            rx = __dadd_rn(cx, __dmul_rn(ux, rr));
            ry = __dadd_rn(cy, __dmul_rn(uy, rr));
        }
    }
    a.robot[SNP_ROBOT_PX * E + env] = (T)rx;
    a.robot[SNP_ROBOT_PY * E + env] = (T)ry;
}

template <typename T> int launch_push(const snp_crowd *c, cudaStream_t st) {
    PushArgs<T> a;
    a.E = c->E; a.N = c->N; a.W = c->W; a.S = c->W > 0 ? c->S : 0; a.walls_per_env = c->walls_per_env;
    a.EN = (long long)c->E * c->N;
    a.dyn = (const T *)c->dyn; a.stat = (const T *)c->stat; a.walls = (const T *)c->walls; a.robot = (T *)c->robot;
    k_robot_push_out<T><<<(unsigned)((c->E + 127) / 128), 128, 0, st>>>(a);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // namespace
}  // namespace snp

using namespace snp;

extern "C" int snp_robot_push_out(const snp_crowd *c, void *stream) {
    if (!c) { set_error("snp_robot_push_out: null crowd"); return SNP_ERR_INVALID; }
    if (c->E <= 0 || c->N <= 0) { set_error("snp_robot_push_out: E and N must be positive"); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->robot) { set_error("snp_robot_push_out: dyn, stat and robot must be device pointers"); return SNP_ERR_INVALID; }
    if (c->W < 0 || (c->W > 0 && (!c->walls || c->S <= 0))) { set_error("snp_robot_push_out: W=%d but no segment array", c->W); return SNP_ERR_INVALID; }
    if (c->dtype == SNP_F64) return launch_push<double>(c, (cudaStream_t)stream);
    if (c->dtype == SNP_F32) return launch_push<float>(c, (cudaStream_t)stream);
    set_error("dtype %d is neither SNP_F32 nor SNP_F64", c->dtype);
    return SNP_ERR_INVALID;
}
