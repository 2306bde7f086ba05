// snp_step_small.inl -- fused small-crowd step (compiled once per arithmetic type by snp_step_small_f32.cu / _f64.cu)
// fused small-crowd step: goal switching, wall closest points, all-pairs social force, desired
// force, HSFM torque + body-frame projection, explicit Euler, robot motion and the collision / goal / reward
// reductions, for `n_substeps` consecutive update_humans calls in ONE launch.  Agent state makes one HBM round trip per
// launch: it lives in registers across sub-steps and only (x, y, vx, vy) of each entity is republished to shared memory.
//
// Two mappings of the same body:
//   warp-packed (CTA = false): N <= 32 humans per env; floor(32/N) envs share a warp, one lane per human; the only
//       synchronisation is __syncwarp; flag reductions are segmented warp shuffles.
//   block-packed (BLOCK = true): floor(128/N) envs share a 128-thread CTA regardless of warp boundaries (N = 25: 5 envs on
//       125 of 128 lanes instead of 25 of 32), or one env per CTA for 128 < N <= 512; synchronisation is __syncthreads
//       (two per sub-step), reductions and the pair exchange go through shared memory.
// Reference: social_gym/src/motion_model_manager.py:354-373,424-459 (serial path), src/forces.py, src/forces_parallel.py:184-284,
// social_gym/social_nav_gym.py:227-250 (sub-step loop), social_nav_sim.py:949-1029 and social_nav_gym.py:107-118 (checks).
#include <cstdio>
#include "snp_kernels.cuh"

namespace snp {

namespace {

// Resident CTAs per SM the compiler must allow for (register cap = 65536 / (128 threads * min blocks)); tuned on B200, see
// profiles/.  Overridable at build time for experiments.
#ifndef SNP_MINB_F32
#define SNP_MINB_F32 7  // 72 registers: 28 warps per SM, i.e. the 4096 warps of the 4096 x 25 workload are resident in ONE wave
#endif
#ifndef SNP_MINB_F64
#define SNP_MINB_F64 4
#endif

// Pair evaluations of the halved loop issued together (independent, branch-free dependency chains per warp; see halved_rounds).
#ifndef SNP_PAIR_UNROLL
#define SNP_PAIR_UNROLL 3  // measured on B200 (4096 x 25): fp64 2 / 3 / 4 -> 0.229 / 0.2215 / 0.2216 ms, fp32 at 72 registers 0.129 / 0.126 / -
#endif

constexpr int kPairUnroll = SNP_PAIR_UNROLL;
#ifndef SNP_WPB
#define SNP_WPB 4
#endif
constexpr int kWarpsPerBlock = SNP_WPB;  // warp-packed mapping: warps (= independent env groups) per CTA
constexpr int kRobotParamWords = 24;  // Params<T> of the robot staged in shared memory (21 values, padded)
// Entity slots of one env group in shared memory.  Warp-packed mapping: [-N, 0) a second copy of the humans, [0, N) the humans,
// [N] the robot -- so that the partner (i + k) mod N of the halved pair loop is simply slot i - N + k (a pointer that only ever
// increments; no wrap test, no index arithmetic per pair).  Block-packed mapping: [0, N] only.
__host__ __device__ inline int group_stride(int N, bool cta) { return cta ? N + 1 : 2 * N + 1; }
__host__ __device__ inline int slots_per_warp(int N, int epw) { return (epw * (2 * N + 1) + 3) & ~3; }
constexpr int kGoalCache = 4;  // goal lists up to this length are staged in shared memory (goal switches then cost no global load)

// Byte offsets of the dynamic shared-memory regions (same arithmetic on host and device).
template <typename T> struct SmemLayout {
    size_t segs, seg_cnt, ents, rs, red, xchg, tflag, rb, goals, total;
    int slots;
    __host__ __device__ SmemLayout(int nseg, int seg_groups, int W, int slots_, int red_doubles, int xchg_vec2 = 0, int groups = 0,
                                   int robot_groups = 0, int goal_vec2 = 0) : slots(slots_) {
        size_t off = sizeof(T) == 8 ? kExpN * sizeof(double) : 0;  // exp table (fp64 only)
        segs = off; off += sizeof(Seg<T>) * (size_t)nseg * seg_groups;
        off = (off + 15) & ~size_t(15);
        seg_cnt = off; off += sizeof(int) * (size_t)(W > 0 ? W : 1) * seg_groups;
        off = (off + 31) & ~size_t(31);
        ents = off; off += sizeof(Ent<T>) * (size_t)slots * 2;
        rs = off; off += sizeof(Vec2<T>) * (size_t)slots;   // r + safety at the stride of the entity planes (.a; .b unused)
        off = (off + 15) & ~size_t(15);
        red = off; off += sizeof(double) * (size_t)red_doubles;
        off = (off + 15) & ~size_t(15);
        xchg = off; off += sizeof(Vec2<T>) * (size_t)xchg_vec2;   // block-packed halved pair loop: [rounds][lanes] of (fx, fy)
        tflag = off; off += sizeof(int) * (size_t)groups;
        off = (off + 15) & ~size_t(15);
        rb = off; off += sizeof(T) * (size_t)(robot_groups ? kRobotParamWords + robot_groups * SNP_ROBOT_FIELDS : 0);  // robot_mode 2
        off = (off + 15) & ~size_t(15);
        goals = off; off += sizeof(Vec2<T>) * (size_t)goal_vec2;  // [G][threads] goal lists (G <= kGoalCache)
        total = off + 16;
    }
};

// ---- checks (double arithmetic, formula-exact; see snp_math.cuh) ----

// social_nav_sim.py:986-1029.  Returns info code, writes reward / terminated / truncated.
__device__ __forceinline__ int reward_and_info(bool collision, double dmin, bool goal, double t, const double *c, double &reward,
                                               bool &terminated, bool &truncated) {
    if (t >= __dsub_rn(c[0], 1.0)) { reward = 0.0; truncated = true; terminated = false; return 1; }
    if (collision) { reward = c[1]; truncated = false; terminated = true; return 2; }
    if (goal) { reward = c[2]; truncated = false; terminated = true; return 3; }
    if (dmin < c[3]) { reward = __dmul_rn(__dmul_rn(__dsub_rn(dmin, c[3]), c[4]), c[5]); truncated = false; terminated = false; return 4; }
    reward = 0.0; truncated = false; terminated = false; return 0;
}

// Segmented min over the lanes [gbase, gbase + n) of a warp; result valid in the group's lane 0.
__device__ __forceinline__ double seg_min(double v, int i, int n, unsigned mask) {
    for (int off = 1; off < n; off <<= 1) {
        const double o = __shfl_down_sync(mask, v, off);
        if (i + off < n) v = o < v ? o : v;
    }
    return v;
}

// Segmented max over the lanes [gbase, gbase + n) of a warp, broadcast to every lane of the group.
template <typename T> __device__ __forceinline__ T seg_max_bcast(T v, int i, int n, int gbase, unsigned mask) {
    for (int off = 1; off < n; off <<= 1) {
        const T o = __shfl_down_sync(mask, v, off);
        if (i + off < n) v = o > v ? o : v;
    }
    return __shfl_sync(mask, v, gbase);
}

// One evaluation of the halved loop: lane i against its partner; for Moussaid the lower index is agent 1 (forces.py:145-151).
template <typename T, int SOC, bool CONTACT>
__device__ __forceinline__ T halved_eval(const Params<T> &P, const double *tbl, const Agent<T> &me, const Vec2<T> op, const Vec2<T> ov, T rsj, bool sw,
                                         T &fx, T &fy) {
    if (SOC == 2) {
        const T rd = pair_eval<T, SOC, CONTACT>(P, tbl, sw ? op.a : me.px, sw ? op.b : me.py, sw ? ov.a : me.vx, sw ? ov.b : me.vy, sw ? rsj : me.rs,
                                                sw ? me.px : op.a, sw ? me.py : op.b, sw ? me.vx : ov.a, sw ? me.vy : ov.b, sw ? me.rs : rsj, fx, fy);
        fx = sw ? -fx : fx; fy = sw ? -fy : fy;
        return rd;
    }
    return pair_eval<T, SOC, CONTACT>(P, tbl, me.px, me.py, me.vx, me.vy, me.rs, op.a, op.b, ov.a, ov.b, rsj, fx, fy);
}

// U rounds of the halved loop: partners at pp[0 .. U) (see group_stride: consecutive slots, no wrap), the lanes whose partner is
// me walk downwards from `src` and wrap at the group's first lane.  The U evaluations are independent and branch-free, so their
// dependency chains interleave; ONE vote covers the rare contact re-evaluation of all of them (the partners' velocities are only
// loaded there, except for Moussaid whose law needs them).
template <typename T, int SOC, int U>
__device__ __forceinline__ void halved_rounds(const Params<T> &P, const double *tbl, const Vec2<T> *pp, const Vec2<T> *pv, const Vec2<T> *pr,
                                              const Agent<T> &me, int k0, int wrap_at, int N, unsigned wmask, int gbase, int &src, T &fsx, T &fsy) {
    Vec2<T> op[U], ov[U];
    T rsj[U], fx[U], fy[U];
    int from[U];
    bool contact = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        op[u] = pp[u];
        rsj[u] = pr[u].a;
        ov[u] = SOC == 2 ? pv[u] : Vec2<T>{T(0), T(0)};
        src = (src == gbase) ? gbase + N - 1 : src - 1;  // the lane whose partner in this round is me: (i - k) mod N
        from[u] = src;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        contact |= Real<T>::positive_(halved_eval<T, SOC, false>(P, tbl, me, op[u], ov[u], rsj[u], k0 + u >= wrap_at, fx[u], fy[u]));
    if (__any_sync(wmask, contact)) {
#pragma unroll
        for (int u = 0; u < U; ++u) halved_eval<T, SOC, true>(P, tbl, me, op[u], pv[u], rsj[u], k0 + u >= wrap_at, fx[u], fy[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const T rx = __shfl_sync(wmask, fx[u], from[u]), ry = __shfl_sync(wmask, fy[u], from[u]);
        fsx += fx[u] - rx; fsy += fy[u] - ry;
    }
}

// Social force of one human with every pair of humans evaluated ONCE per warp (newton's third law): in round k lane i
// evaluates the pair {i, (i+k) mod N}, keeps +f and hands -f to the partner's lane through a warp shuffle, so a crowd of N
// needs (N-1)/2 evaluations per lane instead of N-1 (for even N the antipodal pair is evaluated by both ends).  Valid when the
// law is antisymmetric: uniform parameters, and for Moussaid the reference's symmetric path (lower index is agent 1,
// forces.py:145-151).  Accumulation order differs from the reference's j-ascending order (rounding-level effect only).
// pos / vel / rs16: the env group's planes, index 0 = human 0 (copies at [-N, 0), robot at [N]).
template <typename T, int SOC>
__device__ __forceinline__ void social_force_halved(const Params<T> &P, const double *tbl, const Vec2<T> *pos, const Vec2<T> *vel, const Vec2<T> *rs16,
                                                    const Agent<T> &me, int i, int N, bool with_robot, unsigned wmask, int gbase, T &fsx, T &fsy) {
    const Vec2<T> *pp = pos + (i - N), *pv = vel + (i - N), *pr = rs16 + (i - N);  // slot of (i + k) mod N = pp + k
    const int wrap_at = N - i;  // rounds k >= wrap_at pair me with a LOWER index
    const int rounds = (N - 1) >> 1;
    int src = gbase + i;
    int k = 1;
    for (; k + kPairUnroll - 1 <= rounds; k += kPairUnroll)
        halved_rounds<T, SOC, kPairUnroll>(P, tbl, pp + k, pv + k, pr + k, me, k, wrap_at, N, wmask, gbase, src, fsx, fsy);
    // remaining rounds (fewer than kPairUnroll) in flight TOGETHER as well: a crowd of 5 has 2 rounds in all, and a lone warp per
    // SMSP (4096 envs x 5 humans = 683 warps on 592 schedulers) has nothing else to hide their latency with
    if constexpr (kPairUnroll > 2) {
        if (rounds - k + 1 >= 2) {
            halved_rounds<T, SOC, 2>(P, tbl, pp + k, pv + k, pr + k, me, k, wrap_at, N, wmask, gbase, src, fsx, fsy);
            k += 2;
        }
    }
    for (; k <= rounds; ++k) halved_rounds<T, SOC, 1>(P, tbl, pp + k, pv + k, pr + k, me, k, wrap_at, N, wmask, gbase, src, fsx, fsy);
    // tail: the antipodal pair of an even crowd (both ends evaluate it) and the robot, which exerts force but feels none
    // (forces.py:146,151).  Two self-contained blocks (each with its own vote): nothing is zero-initialised or carried for the case
    // that does not occur -- an odd crowd with a robot, the benchmark shape, runs exactly one pair evaluation here.
    if (!(N & 1)) {
        const Vec2<T> oa = pp[k], ova = pv[k];
        const T rsa = pr[k].a;
        const bool sw = k >= wrap_at;
        T fax, fay;
        if (__any_sync(wmask, Real<T>::positive_(halved_eval<T, SOC, false>(P, tbl, me, oa, ova, rsa, sw, fax, fay))))
            halved_eval<T, SOC, true>(P, tbl, me, oa, ova, rsa, sw, fax, fay);
        fsx += fax; fsy += fay;
    }
    if (with_robot) {
        const Vec2<T> orb = pos[N];
        const T rsr = rs16[N].a;
        T frx, fry;
        const bool contact = Real<T>::positive_(pair_eval<T, SOC, false>(P, tbl, me.px, me.py, me.vx, me.vy, me.rs, orb.a, orb.b, SOC == 2 ? vel[N].a : T(0),
                                                                           SOC == 2 ? vel[N].b : T(0), rsr, frx, fry));
        if (__any_sync(wmask, contact)) {
            const Vec2<T> ovr = vel[N];
            pair_eval<T, SOC, true>(P, tbl, me.px, me.py, me.vx, me.vy, me.rs, orb.a, orb.b, ovr.a, ovr.b, rsr, frx, fry);
        }
        fsx += frx; fsy += fry;
    }
}

// ---- robot driven by its own SFM / HSFM model (robot_mode 2; motion_model_manager.py:593-653) ----
// Kept out of line so that the register allocation of the hot human path is untouched; the robot's model is a run-time value
// (it may differ from the humans').  `rb` is the robot's state in shared memory, indexed by SNP_ROBOT_*.

// Force of human (x2, y2, ...) on the robot: the per-agent path compute_social_force_*(index = len(humans)) (mmm:608).
template <typename T>
__device__ __noinline__ void robot_pair(int soc, const Params<T> *RP, const double *tbl, unsigned mask, const T *rb, T x2, T y2, T vx2, T vy2,
                                        T rs2, T *fx, T *fy) {
    const T rx = rb[SNP_ROBOT_PX], ry = rb[SNP_ROBOT_PY], rvx = rb[SNP_ROBOT_VX], rvy = rb[SNP_ROBOT_VY];
    const T rrs = rb[SNP_ROBOT_R] + rb[SNP_ROBOT_SAFETY];
    if (soc == 0) pair_force<T, 0>(*RP, tbl, mask, rx, ry, rvx, rvy, rrs, x2, y2, vx2, vy2, rs2, *fx, *fy);
    else if (soc == 1) pair_force<T, 1>(*RP, tbl, mask, rx, ry, rvx, rvy, rrs, x2, y2, vx2, vy2, rs2, *fx, *fy);
    else pair_force<T, 2>(*RP, tbl, mask, rx, ry, rvx, rvy, rrs, x2, y2, vx2, vy2, rs2, *fx, *fy);
}

// compute_robot_forces + Euler (mmm:593-629) by ONE lane, given the summed force of the humans (fsx, fsy).
template <typename T>
__device__ __noinline__ void robot_update(int rtype, const Params<T> *RPp, const double *tbl, const Seg<T> *segs, const int *seg_cnt, int W, int S,
                                          T *rb, T fsx, T fsy, T dt, bool just_velocities) {
    using R = Real<T>;
    const Params<T> &RP = *RPp;
    const int obs = (rtype == 1 || rtype == 4 || rtype == 7) ? 1 : 0, headed = rtype / 3;
    const unsigned self = 1u << (threadIdx.x & 31);
    Agent<T> m;
    m.px = rb[SNP_ROBOT_PX]; m.py = rb[SNP_ROBOT_PY]; m.vx = rb[SNP_ROBOT_VX]; m.vy = rb[SNP_ROBOT_VY]; m.th = rb[SNP_ROBOT_TH];
    m.bvx = rb[SNP_ROBOT_BVX]; m.bvy = rb[SNP_ROBOT_BVY]; m.om = rb[SNP_ROBOT_OM]; m.dfx = rb[SNP_ROBOT_DFX]; m.dfy = rb[SNP_ROBOT_DFY];
    m.r = rb[SNP_ROBOT_R]; m.m = rb[SNP_ROBOT_M]; m.vd = rb[SNP_ROBOT_VD]; m.rs = m.r + rb[SNP_ROBOT_SAFETY];
    m.gx = rb[SNP_ROBOT_GX]; m.gy = rb[SNP_ROBOT_GY];
    agent_static<T>(RP, m);
    if (np_norm(m.gx - m.px, m.gy - m.py) < m.r && rb[SNP_ROBOT_GCNT] > T(1.5)) {  // update_goals(robot) (mmm:598): rotate the 2-goal list
        const T ox = m.gx, oy = m.gy;
        m.gx = rb[SNP_ROBOT_GX2]; m.gy = rb[SNP_ROBOT_GY2];
        rb[SNP_ROBOT_GX2] = ox; rb[SNP_ROBOT_GY2] = oy;
    }
    m.cs = T(1); m.sn = T(0);
    if (headed) R::sincos_(m.th, &m.sn, &m.cs);  // v = R(yaw) bv was refreshed when the state was published (mmm:605)
    T fox = T(0), foy = T(0);
    if (W > 0) {
        if (obs == 0) obstacle_force<T, 0>(RP, tbl, self, segs, seg_cnt, W, S, false, m.px, m.py, m.vx, m.vy, m.rs, fox, foy);
        else obstacle_force<T, 1>(RP, tbl, self, segs, seg_cnt, W, S, false, m.px, m.py, m.vx, m.vy, m.rs, fox, foy);
    }
    desired_force<T>(RP, m, false);
    const T px0 = m.px, py0 = m.py, th0 = m.th, cs0 = m.cs, sn0 = m.sn;
    if (headed == 0) integrate<T, 0>(RP, m, fox, foy, fsx, fsy, dt);
    else if (headed == 1) integrate<T, 1>(RP, m, fox, foy, fsx, fsy, dt);
    else integrate<T, 2>(RP, m, fox, foy, fsx, fsy, dt);
    if (just_velocities) {  // mmm:73,79-81: position and yaw stay; v = R(yaw) bv with the UNCHANGED yaw (mmm:85)
        m.px = px0; m.py = py0;
        if (headed) {
            m.th = th0;
            m.vx = np_mv(cs0, -sn0, m.bvx, m.bvy);
            m.vy = np_mv(sn0, cs0, m.bvx, m.bvy);
        }
    }
    rb[SNP_ROBOT_PX] = m.px; rb[SNP_ROBOT_PY] = m.py; rb[SNP_ROBOT_VX] = m.vx; rb[SNP_ROBOT_VY] = m.vy; rb[SNP_ROBOT_TH] = m.th;
    rb[SNP_ROBOT_BVX] = m.bvx; rb[SNP_ROBOT_BVY] = m.bvy; rb[SNP_ROBOT_OM] = m.om; rb[SNP_ROBOT_DFX] = m.dfx; rb[SNP_ROBOT_DFY] = m.dfy;
    rb[SNP_ROBOT_GX] = m.gx; rb[SNP_ROBOT_GY] = m.gy;
}

// RobotAgent.step with unicycle kinematics (robot_agent.py:116-136), action = ActionRot(v, r): the position advances along
// yaw + r, the yaw becomes (yaw + r) % 2 pi -- r is added at EVERY call, whatever delta_t -- and the velocity points along the new
// yaw.  Double arithmetic (what the reference computes in); by the group's leader lane, out of line.
template <typename T> __device__ __noinline__ void robot_unicycle_step(T *rb, double v, double r, double dt) {
    const double two_pi = 6.283185307179586;
    const double yaw = (double)rb[SNP_ROBOT_TH] + r;
    double sn, cs;
    sincos(yaw, &sn, &cs);
    rb[SNP_ROBOT_PX] = (T)__dadd_rn((double)rb[SNP_ROBOT_PX], __dmul_rn(__dmul_rn(cs, v), dt));
    rb[SNP_ROBOT_PY] = (T)__dadd_rn((double)rb[SNP_ROBOT_PY], __dmul_rn(__dmul_rn(sn, v), dt));
    double w = fmod(yaw, two_pi);   // Python's float %: the result takes the divisor's sign
    if (w < 0.0) w += two_pi;
    sincos(w, &sn, &cs);
    rb[SNP_ROBOT_TH] = (T)w;
    rb[SNP_ROBOT_VX] = (T)__dmul_rn(cs, v); rb[SNP_ROBOT_VY] = (T)__dmul_rn(sn, v);
}

template <typename T> __device__ __forceinline__ T seg_sum(T v, int i, int n, unsigned mask) {  // valid in the group's lane 0
    for (int off = 1; off < n; off <<= 1) {
        const T o = __shfl_down_sync(mask, v, off);
        if (i + off < n) v += o;
    }
    return v;
}

template <typename T, int SOC, int OBS, int HEADED, bool CTA, bool PER_AGENT, bool HALF, bool ROBOT2>
__global__ void __launch_bounds__(CTA ? 512 : kWarpsPerBlock * 32, CTA ? 1 : (sizeof(T) == 4 ? SNP_MINB_F32 : SNP_MINB_F64)) k_step(const KArgs<T> a) {
    // CTA == true is the block-packed mapping (a.gpb env groups per CTA), CTA == false the warp-packed one (a.epw per warp).
    using R = Real<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int N = a.N;
    const int lane = threadIdx.x & 31;
    // ---- thread -> (env, human) ----
    int i, g;            // human index in env, group slot in the block
    long long env;
    bool live;
    unsigned wmask = 0xffffffffu;
    if constexpr (CTA) {
        g = threadIdx.x / N;
        i = threadIdx.x - g * N;
        env = (long long)blockIdx.x * a.gpb + g;
        live = g < a.gpb && env < a.E;
        wmask = __ballot_sync(0xffffffffu, live);  // lanes of this warp that take part in votes
    } else {
        const int warp = threadIdx.x >> 5;
        const int gw = lane / N;  // group within the warp
        i = lane - gw * N;
        g = warp * a.epw + gw;
        env = ((long long)blockIdx.x * kWarpsPerBlock + warp) * a.epw + gw;
        live = gw < a.epw && env < a.E;
        wmask = __ballot_sync(0xffffffffu, live);
    }
    const int M = N + (a.consider_robot ? 1 : 0);  // entities exerting force
    const int groups = CTA ? a.gpb : kWarpsPerBlock * a.epw;
    const int rounds = (N - 1) >> 1;  // halved pair loop

    // ---- shared memory carve-up: [exp table][segments][segment counts][entities x2][r+safety][reduction scratch] ----
    const int nseg = a.W * a.S;
    const int seg_groups = a.walls_per_env ? groups : 1;
    const int spw = slots_per_warp(N, a.epw);  // warp-packed: entity slots per warp
    const bool gcache = a.G <= kGoalCache;     // goal lists staged in shared memory
    const SmemLayout<T> lay(nseg, seg_groups, a.W, CTA ? a.gpb * (N + 1) : kWarpsPerBlock * spw, CTA ? a.gpb * N : 0,
                            (CTA && HALF) ? rounds * a.gpb * N : 0, CTA ? a.gpb : 0, ROBOT2 ? groups : 0, gcache ? a.G * (int)blockDim.x : 0);
    double *exp_tbl_s = reinterpret_cast<double *>(smem_raw);
    Seg<T> *segs_all = reinterpret_cast<Seg<T> *>(smem_raw + lay.segs);
    int *seg_cnt_all = reinterpret_cast<int *>(smem_raw + lay.seg_cnt);
    const int slots = lay.slots;
    Vec2<T> *ents0 = reinterpret_cast<Vec2<T> *>(smem_raw + lay.ents);  // [2 buffers][pos | vel][slots]
    Vec2<T> *rs_all = reinterpret_cast<Vec2<T> *>(smem_raw + lay.rs);
    Vec2<T> *goal_s = reinterpret_cast<Vec2<T> *>(smem_raw + lay.goals) + threadIdx.x;  // this thread's goal k at goal_s[k * blockDim.x]
    double *red = reinterpret_cast<double *>(smem_raw + lay.red) + (CTA ? g * N : 0);  // block-packed only: this group's [N] doubles
    Vec2<T> *xchg = reinterpret_cast<Vec2<T> *>(smem_raw + lay.xchg) + (CTA ? g * N : 0);
    int *tflag = reinterpret_cast<int *>(smem_raw + lay.tflag);
    Params<T> *RPs = reinterpret_cast<Params<T> *>(smem_raw + lay.rb);  // robot_mode 2 only
    T *rb = reinterpret_cast<T *>(smem_raw + lay.rb) + kRobotParamWords + (size_t)g * SNP_ROBOT_FIELDS;

    if (sizeof(T) == 8) exp_table_init(exp_tbl_s);
    if constexpr (ROBOT2) { if (threadIdx.x == 0) *RPs = a.RP; }
    // walls -> shared memory (whole block cooperates, before anyone leaves)
    if (nseg > 0) {
        const int total = nseg * seg_groups;
        for (int k = threadIdx.x; k < total; k += blockDim.x) {
            const int sg = k / nseg, s = k - sg * nseg;
            long long wenv = 0;
            if (a.walls_per_env) {
                wenv = CTA ? ((long long)blockIdx.x * a.gpb + sg) : ((long long)blockIdx.x * kWarpsPerBlock * a.epw + sg);
                if (wenv >= a.E) wenv = a.E - 1;
            }
            const T *w = a.walls + ((size_t)wenv * nseg + s) * 4;
            segs_all[k] = make_seg<T>(w[0], w[1], w[2], w[3]);
        }
        __syncthreads();
        // valid slots per polygon: NaN padding is a suffix of each polygon's slots (motion_model_manager.py:270-275)
        for (int k = threadIdx.x; k < a.W * seg_groups; k += blockDim.x) {
            int c = 0;
            while (c < a.S && segs_all[(size_t)k * a.S + c].ax == segs_all[(size_t)k * a.S + c].ax) ++c;
            seg_cnt_all[k] = c;
        }
    }
    __syncthreads();
    if constexpr (!CTA) { if (!live) return; }

    // slot of this group's human 0 (see group_stride)
    const int gslot = CTA ? g * (N + 1) : ((threadIdx.x >> 5) * spw + (g - (threadIdx.x >> 5) * a.epw) * (2 * N + 1) + N);
    const Seg<T> *segs = segs_all + (a.walls_per_env ? (size_t)g * nseg : 0);
    const int *seg_cnt = seg_cnt_all + (a.walls_per_env ? g * a.W : 0);
    Vec2<T> *rs_g = rs_all + gslot;
    const bool leader = live && i == 0;
    const bool has_robot = a.robot != nullptr;
    const long long EN = a.EN;
    const long long aidx = live ? env * N + i : 0;

    // ---- load ----
    Agent<T> me;
    Params<T> P = a.P;
    int gidx = 0, gcnt = 1;
    bool goals_differ = true;  // false: every entry of the goal list is the same point (the zero-speed "static obstacle" humans of the
                               // reference's CCSO scenario have goals [p, p]): rotating such a list changes nothing, so the switch is skipped
    if (live) {
        me.px = a.dyn[SNP_DYN_PX * EN + aidx]; me.py = a.dyn[SNP_DYN_PY * EN + aidx];
        me.vx = a.dyn[SNP_DYN_VX * EN + aidx]; me.vy = a.dyn[SNP_DYN_VY * EN + aidx];
        me.dfx = a.dyn[SNP_DYN_DFX * EN + aidx]; me.dfy = a.dyn[SNP_DYN_DFY * EN + aidx];
        if (HEADED) {
            me.th = a.dyn[SNP_DYN_TH * EN + aidx];
            me.bvx = a.dyn[SNP_DYN_BVX * EN + aidx]; me.bvy = a.dyn[SNP_DYN_BVY * EN + aidx];
            me.om = a.dyn[SNP_DYN_OM * EN + aidx];
        } else { me.th = me.bvx = me.bvy = me.om = T(0); }
        me.r = a.stat[SNP_STAT_R * EN + aidx]; me.m = a.stat[SNP_STAT_M * EN + aidx];
        me.vd = a.stat[SNP_STAT_VD * EN + aidx];
        me.rs = me.r + a.stat[SNP_STAT_SAFETY * EN + aidx];
        gidx = a.goal_idx[aidx]; gcnt = a.goal_cnt[aidx];
        me.gx = a.goals[((size_t)gidx * 2 + 0) * EN + aidx]; me.gy = a.goals[((size_t)gidx * 2 + 1) * EN + aidx];
        if constexpr (PER_AGENT) {
            double p[20];
#pragma unroll
            for (int k = 0; k < 20; ++k) p[k] = (double)a.agent_params[(size_t)k * EN + aidx];
            P = make_params<T>(p);
        }
        agent_static<T>(P, me);
        rs_g[i].a = me.rs;
        if constexpr (!CTA) rs_g[i - N].a = me.rs;
        if (gcache) {
            goals_differ = false;
            for (int k = 0; k < a.G; ++k) {
                const Vec2<T> gk{a.goals[((size_t)k * 2 + 0) * EN + aidx], a.goals[((size_t)k * 2 + 1) * EN + aidx]};
                goal_s[(size_t)k * blockDim.x] = gk;
                if (k < gcnt && (gk.a != me.gx || gk.b != me.gy)) goals_differ = true;
            }
        }
    } else {
        me = Agent<T>{};
    }
    me.cs = T(1); me.sn = T(0);
    // robot (every lane keeps a copy; the group leader owns the shared-memory entity and the write-back)
    T rpx = T(0), rpy = T(0), rvx = T(0), rvy = T(0), rr = T(0), rrs = T(0), rgx = T(0), rgy = T(0), ax = T(0), ay = T(0);
    if (has_robot && live) {
        const long long E = a.E;
        rpx = a.robot[SNP_ROBOT_PX * E + env]; rpy = a.robot[SNP_ROBOT_PY * E + env];
        rvx = a.robot[SNP_ROBOT_VX * E + env]; rvy = a.robot[SNP_ROBOT_VY * E + env];
        rr = a.robot[SNP_ROBOT_R * E + env]; rrs = rr + a.robot[SNP_ROBOT_SAFETY * E + env];
        rgx = a.robot[SNP_ROBOT_GX * E + env]; rgy = a.robot[SNP_ROBOT_GY * E + env];
        if (a.action) { ax = a.action[env]; ay = a.action[E + env]; }
        if (leader) rs_g[N].a = rrs;
        if constexpr (ROBOT2) if (leader) {  // the robot's full state lives in shared memory (one lane updates it)
#pragma unroll
            for (int f = 0; f < SNP_ROBOT_FIELDS; ++f) rb[f] = a.robot[(size_t)f * E + env];
            if (a.robot_mode == 2 && a.robot_type >= 3 && a.robot_every <= 1) {  // headed robot: linear velocity = R(yaw) bv (mmm:605); with robot_every > 1
                T sn, cs;                                   // the velocity of the last refresh is what moves the pose (mmm:655-657)
                R::sincos_(rb[SNP_ROBOT_TH], &sn, &cs);
                rb[SNP_ROBOT_VX] = np_mv(cs, -sn, rb[SNP_ROBOT_BVX], rb[SNP_ROBOT_BVY]);
                rb[SNP_ROBOT_VY] = np_mv(sn, cs, rb[SNP_ROBOT_BVX], rb[SNP_ROBOT_BVY]);
            }
        }
    }
    double tnow = (a.time_now && live) ? a.time_now[env] : 0.0;

    int flags = 0;
    double out_dmin = 0.0, out_reward = 0.0, out_admin = 0.0;

    // ---- pre-step checks: swept collision + goal + reward (gym:232-234) ----
    // unicycle action (v, r): the swept test and the goal test use the velocity v (cos, sin)(yaw + r) (sim:973, robot_agent.py:122)
    // (ROBOT2 instantiations only: they also carry the unicycle robot, so that the common kernel has none of its code)
    if (ROBOT2 && a.robot_mode == 3 && has_robot && live) {
        const double uv = (double)ax, ur = (double)ay;
        double sn_, cs_;
        sincos((double)a.robot[SNP_ROBOT_TH * a.E + env] + ur, &sn_, &cs_);
        ax = (T)__dmul_rn(uv, cs_); ay = (T)__dmul_rn(uv, sn_);
    }
    if (a.pre_checks) {
        double cd = CUDART_INF;
        if (live) cd = swept_distance((double)me.px, (double)me.py, (double)me.vx, (double)me.vy, (double)me.r, (double)rpx, (double)rpy,
                                      (double)rr, (double)ax, (double)ay, a.consts[5]);
        bool collision; double dmin;
        if constexpr (CTA) {
            if (live) red[i] = cd;
            __syncthreads();
            collision = false; dmin = CUDART_INF;
            if (leader) for (int k = 0; k < N; ++k) { const double c = red[k]; if (c < 0) { collision = true; break; } else if (c < dmin) dmin = c; }
            __syncthreads();
        } else {
            const int gbase = lane - i;
            const unsigned gm = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << gbase;
            const unsigned hit = __ballot_sync(wmask, cd < 0) & gm;
            const int first = hit ? (__ffs(hit) - 1 - gbase) : N;
            collision = hit != 0;
            const double v = (i < first && cd >= 0) ? cd : CUDART_INF;
            dmin = seg_min(v, i, N, wmask);
        }
        if (leader) {
            const double endx = __dadd_rn((double)rpx, __dmul_rn((double)ax, a.consts[5]));
            const double endy = __dadd_rn((double)rpy, __dmul_rn((double)ay, a.consts[5]));
            const bool goal = xnorm_np(__dsub_rn(endx, (double)rgx), __dsub_rn(endy, (double)rgy)) < (double)rr;
            bool term, trunc;
            const int code = reward_and_info(collision, dmin, goal, tnow, a.consts, out_reward, term, trunc);
            out_dmin = dmin;
            flags |= (collision ? SNP_FLAG_COLLISION : 0) | (goal ? SNP_FLAG_REACHING_GOAL : 0) | (term ? SNP_FLAG_TERMINATED : 0) |
                     (trunc ? SNP_FLAG_TRUNCATED : 0) | (code << SNP_FLAG_INFO_SHIFT);
        }
    }

    if (HEADED && live && a.n_substeps > 0) {  // mmm:448 / fp:254-256: v = R(yaw) bv before any force is evaluated
        R::sincos_(me.th, &me.sn, &me.cs);
        me.vx = np_mv(me.cs, -me.sn, me.bvx, me.bvy);
        me.vy = np_mv(me.sn, me.cs, me.bvx, me.bvy);
    }

    // ---- fused sub-steps ----
    bool touched = false;
    const T dt = a.dt;
    // every squared distance that can round to a touching distance lies below (r + r_robot)^2 (1 + 2^-48)
    const double touch_thr = __dadd_rn((double)me.r, (double)rr);
    const double touch_hi = touch_thr * touch_thr * (1.0 + 3.5527136788005009e-15);
    // loop-invariant switches as predicates (otherwise every sub-step re-reads them from the constant bank)
    const bool move_robot = a.robot_mode == 1 && has_robot, robot_seen = a.consider_robot != 0, walls_on = a.W > 0, numba_sem = a.numba != 0;
    const bool touch_on = a.track_touch && has_robot, clock_on = a.time_now != nullptr;
    const int n_sub = a.n_substeps;
    for (int s = 0; s < n_sub; ++s) {
        const EntView<T> ents{ents0 + (size_t)(s & 1) * 2 * slots + gslot, ents0 + (size_t)(s & 1) * 2 * slots + slots + gslot};
        if constexpr (!CTA) { if (live) ents.put(i - N, me.px, me.py, me.vx, me.vy); }  // second copy: the halved loop's wrap-free run
        if (move_robot) {  // robot_agent.py:126-131 (holonomic): p = p + a*dt ; v = a
            if (sizeof(T) == 8) { rpx = (T)__dadd_rn((double)rpx, __dmul_rn((double)ax, (double)dt)); rpy = (T)__dadd_rn((double)rpy, __dmul_rn((double)ay, (double)dt)); }
            else { rpx = rpx + ax * dt; rpy = rpy + ay * dt; }
            rvx = ax; rvy = ay;
        }
        if constexpr (ROBOT2) {
            if (a.robot_mode == 3 && has_robot) {  // unicycle robot: uniform branch; the pose lives in the group's shared record
                if (leader) robot_unicycle_step<T>(rb, (double)a.action[env], (double)a.action[a.E + env], (double)dt);  // (v, r)
                if constexpr (CTA) __syncthreads(); else __syncwarp(wmask);
                rpx = rb[SNP_ROBOT_PX]; rpy = rb[SNP_ROBOT_PY]; rvx = rb[SNP_ROBOT_VX]; rvy = rb[SNP_ROBOT_VY];
                if (leader && a.consider_robot) ents.put(N, rpx, rpy, rvx, rvy);
            }
        }
        if (live) ents.put(i, me.px, me.py, me.vx, me.vy);
        if (!ROBOT2 && leader && robot_seen) ents.put(N, rpx, rpy, rvx, rvy);
        if constexpr (CTA) __syncthreads(); else __syncwarp(wmask);

        if constexpr (ROBOT2) if (a.robot_mode == 2) {
            // update_robot first (gym:262): every human lane evaluates its force on the robot with the ROBOT's model and
            // parameters, the group sums them, the leader integrates the robot and republishes it; the humans then see the
            // moved robot (gym:264).
            // a.robot_every >= 1 -- SocialNavSim.update / control_robot (sim:476-529) instead: the humans see the robot's state from
            // BEFORE its update (sim:484-491), so that state is what goes into entity slot N; with robot_every > 1 the pose advances
            // every sub-step with the last velocity (update_robot_pose, mmm:655-657, yaw unwrapped) and only every robot_every-th
            // sub-step refreshes the velocities (update_robot just_velocities with dt = ROBOT_SAMPLING_TIME, sim:523-524).
            const bool sched = a.robot_every >= 1, jv = a.robot_every > 1;
            const bool upd = !jv || ((a.robot_phase + s) % a.robot_every) == 0;  // uniform
            if (sched) {
                if (leader) {
                    if (a.consider_robot) ents.put(N, rb[SNP_ROBOT_PX], rb[SNP_ROBOT_PY], rb[SNP_ROBOT_VX], rb[SNP_ROBOT_VY]);
                    if (jv) {
                        rb[SNP_ROBOT_PX] = fma_<T>(rb[SNP_ROBOT_VX], dt, rb[SNP_ROBOT_PX]);
                        rb[SNP_ROBOT_PY] = fma_<T>(rb[SNP_ROBOT_VY], dt, rb[SNP_ROBOT_PY]);
                        rb[SNP_ROBOT_TH] = fma_<T>(rb[SNP_ROBOT_OM], dt, rb[SNP_ROBOT_TH]);
                        if (upd && a.robot_type >= 3) {  // compute_robot_forces refreshes v = R(yaw) bv before any force (mmm:605)
                            T sn_, cs_;
                            R::sincos_(rb[SNP_ROBOT_TH], &sn_, &cs_);
                            const T bx = rb[SNP_ROBOT_BVX], by = rb[SNP_ROBOT_BVY];
                            rb[SNP_ROBOT_VX] = np_mv(cs_, -sn_, bx, by);
                            rb[SNP_ROBOT_VY] = np_mv(sn_, cs_, bx, by);
                        }
                    }
                }
                if constexpr (CTA) __syncthreads(); else __syncwarp(wmask);
            }
            T rfx = T(0), rfy = T(0);
            if (live && upd) robot_pair<T>(a.robot_type % 3, RPs, exp_tbl_s, wmask, rb, me.px, me.py, me.vx, me.vy, me.rs, &rfx, &rfy);
            if constexpr (CTA) {
                if (live) { red[i] = (double)rfx; }
                __syncthreads();
                double sx = 0.0;
                if (leader) for (int k = 0; k < N; ++k) sx += red[k];
                __syncthreads();
                if (live) { red[i] = (double)rfy; }
                __syncthreads();
                double sy = 0.0;
                if (leader) for (int k = 0; k < N; ++k) sy += red[k];
                rfx = (T)sx; rfy = (T)sy;
            } else {
                rfx = seg_sum<T>(rfx, i, N, wmask); rfy = seg_sum<T>(rfy, i, N, wmask);
            }
            if (leader) {
                if (upd) robot_update<T>(a.robot_type, RPs, exp_tbl_s, segs, seg_cnt, a.W, a.S, rb, rfx, rfy, jv ? a.robot_dt : dt, jv);
                if (a.consider_robot && !sched) ents.put(N, rb[SNP_ROBOT_PX], rb[SNP_ROBOT_PY], rb[SNP_ROBOT_VX], rb[SNP_ROBOT_VY]);
            }
            if constexpr (CTA) __syncthreads(); else __syncwarp(wmask);
            rpx = rb[SNP_ROBOT_PX]; rpy = rb[SNP_ROBOT_PY]; rvx = rb[SNP_ROBOT_VX]; rvy = rb[SNP_ROBOT_VY];
            rgx = rb[SNP_ROBOT_GX]; rgy = rb[SNP_ROBOT_GY];
        }

        T fox = T(0), foy = T(0), fsx = T(0), fsy = T(0);
        if (live) {
            // goal switching (mmm:66-70 '<', fp:226 '<=') and the desired force (forces.py:9-16) share the goal vector: both are
            // evaluated at the position the sub-step starts from
            {
                GoalVec<T> gv = goal_vec<T>(me);
                if (goals_differ && (numba_sem ? (gv.dist <= me.r) : (gv.dist < me.r))) {
                    gidx = (gidx + 1 >= gcnt) ? 0 : gidx + 1;
                    T ngx, ngy;
                    if (gcache) { const Vec2<T> ng = goal_s[(size_t)gidx * blockDim.x]; ngx = ng.a; ngy = ng.b; }
                    else { ngx = a.goals[((size_t)gidx * 2 + 0) * EN + aidx]; ngy = a.goals[((size_t)gidx * 2 + 1) * EN + aidx]; }
                    if (ngx != me.gx || ngy != me.gy) { me.gx = ngx; me.gy = ngy; gv = goal_vec<T>(me); }
                }
                desired_force<T>(P, me, gv, numba_sem);
            }
            // wall force
            if (walls_on) obstacle_force<T, OBS>(P, exp_tbl_s, wmask, segs, seg_cnt, a.W, a.S, numba_sem, me.px, me.py, me.vx, me.vy, me.rs, fox, foy);
            if constexpr (HALF && !CTA) {
                social_force_halved<T, SOC>(P, exp_tbl_s, ents.pos, ents.vel, rs_g, me, i, N, robot_seen, wmask, lane - i, fsx, fsy);
            } else if constexpr (HALF && CTA) {
                // Block-packed halved loop, phase 1: lane i evaluates the pairs {i, (i+k) mod N}, k = 1..rounds, keeps +f and
                // leaves -f for the partner in plane k of the exchange buffer (written at the PARTNER's slot: conflict-free).
                const int L = a.gpb * N;
                int pidx = i;
#pragma unroll 2
                for (int k = 0; k < rounds; ++k) {
                    pidx = (pidx + 1 == N) ? 0 : pidx + 1;
                    const Ent<T> o = ents.get(pidx);
                    const T rsj = rs_g[pidx].a;
                    T fx, fy;
                    if (SOC == 2) {
                        const bool sw = pidx < i;  // the lower index is agent 1 (forces.py:145-151)
                        pair_force<T, SOC>(P, exp_tbl_s, wmask, sw ? o.x : me.px, sw ? o.y : me.py, sw ? o.vx : me.vx, sw ? o.vy : me.vy, sw ? rsj : me.rs,
                                           sw ? me.px : o.x, sw ? me.py : o.y, sw ? me.vx : o.vx, sw ? me.vy : o.vy, sw ? me.rs : rsj, fx, fy);
                        fx = sw ? -fx : fx; fy = sw ? -fy : fy;
                    } else {
                        pair_force<T, SOC>(P, exp_tbl_s, wmask, me.px, me.py, me.vx, me.vy, me.rs, o.x, o.y, o.vx, o.vy, rsj, fx, fy);
                    }
                    fsx += fx; fsy += fy;
                    xchg[(size_t)k * L + pidx] = Vec2<T>{fx, fy};
                }
                if (!(N & 1)) {  // antipodal pair: both ends evaluate it
                    pidx = (pidx + 1 == N) ? 0 : pidx + 1;
                    const Ent<T> o = ents.get(pidx);
                    const T rsj = rs_g[pidx].a;
                    T fx, fy;
                    if (SOC == 2) {
                        const bool sw = pidx < i;
                        pair_force<T, SOC>(P, exp_tbl_s, wmask, sw ? o.x : me.px, sw ? o.y : me.py, sw ? o.vx : me.vx, sw ? o.vy : me.vy, sw ? rsj : me.rs,
                                           sw ? me.px : o.x, sw ? me.py : o.y, sw ? me.vx : o.vx, sw ? me.vy : o.vy, sw ? me.rs : rsj, fx, fy);
                        fx = sw ? -fx : fx; fy = sw ? -fy : fy;
                    } else {
                        pair_force<T, SOC>(P, exp_tbl_s, wmask, me.px, me.py, me.vx, me.vy, me.rs, o.x, o.y, o.vx, o.vy, rsj, fx, fy);
                    }
                    fsx += fx; fsy += fy;
                }
                if (a.consider_robot) {
                    const Ent<T> o = ents.get(N);
                    T fx, fy;
                    pair_force<T, SOC>(P, exp_tbl_s, wmask, me.px, me.py, me.vx, me.vy, me.rs, o.x, o.y, o.vx, o.vy, rs_g[N].a, fx, fy);
                    fsx += fx; fsy += fy;
                }
            } else {
                // j ascending: exactly the accumulation order of forces.py:145-151.  Branch-free body: the self pair contributes
                // an exactly zero force by construction (tiny_ in pair_force), so consecutive pairs interleave in the pipes.
                const bool sym = a.symmetric != 0;
#pragma unroll 2
                for (int j = 0; j < M; ++j) {
                    const Ent<T> o = ents.get(j);
                    const T rsj = rs_g[j].a;
                    T fx, fy;
                    if (SOC == 2) {
                        // symmetric path: the pair is evaluated once with the LOWER index as agent 1 and applied with a minus
                        // sign to the other (forces.py:149-151); only Moussaid's sign(theta) makes that differ from f(i,j).
                        const bool sw = sym && j < i;
                        pair_force<T, SOC>(P, exp_tbl_s, wmask, sw ? o.x : me.px, sw ? o.y : me.py, sw ? o.vx : me.vx, sw ? o.vy : me.vy, sw ? rsj : me.rs,
                                           sw ? me.px : o.x, sw ? me.py : o.y, sw ? me.vx : o.vx, sw ? me.vy : o.vy, sw ? me.rs : rsj, fx, fy);
                        fx = sw ? -fx : fx; fy = sw ? -fy : fy;
                    } else {
                        pair_force<T, SOC>(P, exp_tbl_s, wmask, me.px, me.py, me.vx, me.vy, me.rs, o.x, o.y, o.vx, o.vy, rsj, fx, fy);
                    }
                    fsx += fx; fsy += fy;
                }
            }
        }
        if constexpr (HALF && CTA) {
            // phase 2: one barrier, then every lane collects what its partners left for it (own slot of every plane)
            __syncthreads();
            if (live) {
                const int L = a.gpb * N;
#pragma unroll 4
                for (int k = 0; k < rounds; ++k) {
                    const Vec2<T> r = xchg[(size_t)k * L + i];
                    fsx -= r.a; fsy -= r.b;
                }
            }
        }
        if (live) {
            integrate<T, HEADED>(P, me, fox, foy, fsx, fsy, dt);
        }
        if constexpr (!CTA) {
            // post_update: parallel-traffic respawn (mmm:407-422).  Rare, so the warp first votes; pending humans are then handled
            // one index at a time per env group (the reference's loop is sequential: the max runs over already respawned humans).
            if (a.respawn) {
                const int gbase = lane - i;
                const unsigned gm = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << gbase;
                const bool env_on = !a.respawn_envs || a.respawn_envs[env] != 0;
                unsigned pending = __ballot_sync(wmask, env_on && np_norm(me.px - me.gx, me.py - me.gy) < T(3));
                if (pending) {
                    T rsmax = seg_max_bcast<T>(me.rs, i, N, gbase, wmask);
                    if (a.consider_robot) rsmax = rrs > rsmax ? rrs : rsmax;
                    while (pending) {
                        const unsigned mine = pending & gm;
                        const int r = mine ? (__ffs(mine) - 1 - gbase) : -1;  // lowest pending index of my group
                        T xmax = seg_max_bcast<T>(me.px, i, N, gbase, wmask);
                        if (a.consider_robot) xmax = rpx > xmax ? rpx : xmax;
                        if (i == r) {
                            const T nx = fma_<T>(rsmax, T(2), xmax);
                            me.px = nx > (T)a.respawn_bounds[0] ? nx : (T)a.respawn_bounds[0];
                            const T by = (T)a.respawn_bounds[1];
                            me.py = me.py >= T(0) ? (me.py < by ? me.py : by) : (me.py > -by ? me.py : -by);
                            me.gy = me.py;  // human.set_goals([[goals[0][0], position[1]]])  (mmm:418)
                            gidx = 0; gcnt = 1;
                            const_cast<T *>(a.goals)[(size_t)0 * EN + aidx] = me.gx;
                            const_cast<T *>(a.goals)[(size_t)1 * EN + aidx] = me.gy;
                            const_cast<int *>(a.goal_cnt)[aidx] = 1;
                            if (gcache) goal_s[0] = Vec2<T>{me.gx, me.gy};
                        }
                        unsigned clear = 0;  // every group retires its lowest pending index
                        for (unsigned rest = pending; rest;) {
                            const int b = __ffs(rest) - 1;
                            const int gb = b - (b % N);  // groups start at multiples of N within the warp
                            clear |= 1u << b;
                            rest &= ~(((N >= 32 ? 0xffffffffu : ((1u << N) - 1u))) << gb);
                        }
                        pending &= ~clear;
                    }
                }
            }
        }
        if (live) {
            if (touch_on) {
                // ||p - p_robot|| < r + r_robot (robot_agent.py:37), decided on the square whenever that is safe: the IEEE square
                // root only runs for lanes within 2^-48 (relative) of touching or beyond, which is rare
                const double tx = (double)me.px - (double)rpx, ty = (double)me.py - (double)rpy;
                const double t2 = __fma_rn(ty, ty, __dmul_rn(tx, tx));
                if (t2 < touch_hi) touched |= sqrt(t2) < __dadd_rn((double)me.r, (double)rr);
            }
        }
        if (clock_on) tnow = __dadd_rn(tnow, a.dt_d);
    }

    // ---- post-step checks (gym:107-118) ----
    if (a.post_checks || a.track_touch) {
        double d = 10000.0;
        if (live && a.post_checks) {
            d = __dsub_rn(__dsub_rn(xnorm_np(__dsub_rn((double)me.px, (double)rpx), __dsub_rn((double)me.py, (double)rpy)), (double)me.r), (double)rr);
            if (!(d < 10000.0)) d = 10000.0;
        }
        double admin; bool any_touch;
        if constexpr (CTA) {
            if (threadIdx.x < a.gpb) tflag[threadIdx.x] = 0;
            __syncthreads();
            if (live) red[i] = d;
            if (live && touched) tflag[g] = 1;
            __syncthreads();
            admin = 10000.0;
            if (leader) for (int k = 0; k < N; ++k) admin = red[k] < admin ? red[k] : admin;
            any_touch = leader && tflag[g] != 0;
        } else {
            const int gbase = lane - i;
            const unsigned gm = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << gbase;
            admin = seg_min(d, i, N, wmask);
            any_touch = (__ballot_sync(wmask, touched) & gm) != 0;
        }
        if (leader) {
            if (a.post_checks) {
                const bool agoal = xnorm_np(__dsub_rn((double)rpx, (double)rgx), __dsub_rn((double)rpy, (double)rgy)) < (double)rr;
                out_admin = admin;
                flags |= (admin <= 0.0 ? SNP_FLAG_ACTUAL_COLLISION : 0) | (agoal ? SNP_FLAG_ACTUAL_GOAL : 0);
                if (a.post_checks == 2) {  // imitation_learning_step: reward / info from the ACTUAL end state and end time (gym:269-271)
                    bool term, trunc;
                    const int code = reward_and_info(admin <= 0.0, admin, agoal, tnow, a.consts, out_reward, term, trunc);
                    flags |= (term ? SNP_FLAG_TERMINATED : 0) | (trunc ? SNP_FLAG_TRUNCATED : 0) | (code << SNP_FLAG_INFO_SHIFT);
                }
            }
            if (any_touch) flags |= SNP_FLAG_TOUCHED;
        }
    }

    // ---- store ----
    if (live && a.n_substeps > 0) {
        // peek (get_next_human_observable_states, mmm:691-709): pose and velocities go to a side buffer and the goal index is
        // left alone; the carried desired force is still updated in place -- the reference does not restore it either
        T *out = a.dyn_out ? a.dyn_out : a.dyn;
        out[SNP_DYN_PX * EN + aidx] = me.px; out[SNP_DYN_PY * EN + aidx] = me.py;
        out[SNP_DYN_VX * EN + aidx] = me.vx; out[SNP_DYN_VY * EN + aidx] = me.vy;
        a.dyn[SNP_DYN_DFX * EN + aidx] = me.dfx; a.dyn[SNP_DYN_DFY * EN + aidx] = me.dfy;
        if (HEADED) {
            out[SNP_DYN_TH * EN + aidx] = me.th;
            out[SNP_DYN_BVX * EN + aidx] = me.bvx; out[SNP_DYN_BVY * EN + aidx] = me.bvy;
            out[SNP_DYN_OM * EN + aidx] = me.om;
        }
        if (!a.dyn_out) a.goal_idx[aidx] = gidx;
        else if (a.goal_idx_out) a.goal_idx_out[aidx] = gidx;
        if (a.obs_out) {  // the caller's pinned observation buffer [4][E*N]
            a.obs_out[0 * EN + aidx] = me.px; a.obs_out[1 * EN + aidx] = me.py;
            a.obs_out[2 * EN + aidx] = me.vx; a.obs_out[3 * EN + aidx] = me.vy;
        }
    }
    if (leader) {
        if (has_robot && a.robot_mode == 1) {
            const long long E = a.E;
            a.robot[SNP_ROBOT_PX * E + env] = rpx; a.robot[SNP_ROBOT_PY * E + env] = rpy;
            a.robot[SNP_ROBOT_VX * E + env] = rvx; a.robot[SNP_ROBOT_VY * E + env] = rvy;
        }
        if (ROBOT2 && has_robot && a.n_substeps > 0) {
            const long long E = a.E;
#pragma unroll
            for (int f = 0; f < SNP_ROBOT_FIELDS; ++f)
                if (f != SNP_ROBOT_R && f != SNP_ROBOT_SAFETY && f != SNP_ROBOT_M && f != SNP_ROBOT_VD && f != SNP_ROBOT_GCNT && f != SNP_ROBOT_SPARE)
                    a.robot[(size_t)f * E + env] = rb[f];
        }
        if (a.time_now) a.time_now[env] = tnow;
        if (a.flags) a.flags[env] = flags;
        if (a.flags2) a.flags2[env] = flags;
        if (a.checks) {
            a.checks[env * 4 + 0] = out_dmin; a.checks[env * 4 + 1] = out_reward;
            a.checks[env * 4 + 2] = out_admin; a.checks[env * 4 + 3] = 0.0;
        }
        if (a.checks2) {
            a.checks2[env * 4 + 0] = out_dmin; a.checks2[env * 4 + 1] = out_reward;
            a.checks2[env * 4 + 2] = out_admin; a.checks2[env * 4 + 3] = 0.0;
        }
    }
}

template <typename T, int SOC, int OBS, int HEADED>
int launch_one(const KArgs<T> &a_in, cudaStream_t st) {
    KArgs<T> a = a_in;
    const int nseg = a.W * a.S;
    if (sizeof(T) == 8) SNP_CUDA_OK(ensure_exp_table());
    if (a.N > 512) { set_error("snp_step handles N <= 512 humans per env (got %d); use snp_large_step", a.N); return SNP_ERR_UNSUPPORTED; }
    const bool per_agent = a.agent_params != nullptr;
    // pairs evaluated once per warp / block whenever the law is antisymmetric (see social_force_halved)
    bool half = !per_agent && !a.full_pair_loop && (SOC != 2 || a.symmetric);
    // Mapping: warp-packed for N <= 32 (envs tile a warp; __syncwarp and shuffles only), block-packed above (envs tile a 128-thread
    // CTA).  Block-packing small crowds fills more lanes (N = 25: 125/128 instead of 25/32) but measured SLOWER on B200 (fp64
    // 0.315 vs 0.281 ms, fp32 0.192 vs 0.178 ms per 20-sub-step launch at 4096 x 25): two __syncthreads per sub-step and the
    // shared-memory exchange cost more than the idle lanes.  a.mapping: 0 auto, 1 force warp-packed, 2 force block-packed.
    const int gpb = a.N <= 128 ? 128 / a.N : 1;
    const int block_threads = a.N <= 128 ? 128 : (a.N + 31) / 32 * 32;
    bool cta = a.N > 32;
    if (a.mapping == 1 && a.N <= 32) cta = false;
    if (a.mapping == 2) cta = true;
    if (a.respawn && cta) {  // the parallel-traffic respawn (mmm:407-422) lives in the warp-packed mapping only
        if (a.N <= 32) cta = false;
        else { set_error("parallel-traffic respawn supports at most 32 humans per env (got %d)", a.N); return SNP_ERR_UNSUPPORTED; }
    }
    dim3 grid, block;
    size_t smem;
    if (!cta) {
        a.epw = 32 / a.N;
        const int epb = a.epw * kWarpsPerBlock;
        grid = dim3((unsigned)((a.E + epb - 1) / epb));
        block = dim3(kWarpsPerBlock * 32);
        smem = SmemLayout<T>(nseg, a.walls_per_env ? epb : 1, a.W, kWarpsPerBlock * slots_per_warp(a.N, a.epw), 0, 0, 0, a.robot_mode >= 2 ? epb : 0,
                             a.G <= kGoalCache ? a.G * kWarpsPerBlock * 32 : 0).total;
    } else {
        a.epw = 1;
        a.gpb = gpb;
        grid = dim3((unsigned)((a.E + gpb - 1) / gpb));
        block = dim3((unsigned)block_threads);
        const int rounds = (a.N - 1) >> 1;
        if (half && sizeof(Vec2<T>) * (size_t)rounds * gpb * a.N > 64 * 1024) half = false;  // exchange planes would not fit: ordered loop
        smem = SmemLayout<T>(nseg, a.walls_per_env ? gpb : 1, a.W, gpb * (a.N + 1), gpb * a.N, half ? rounds * gpb * a.N : 0, gpb,
                             a.robot_mode >= 2 ? gpb : 0, a.G <= kGoalCache ? a.G * block_threads : 0).total;
    }
    if (smem > 200 * 1024) { set_error("wall/segment tables need %zu bytes of shared memory (limit 200 KiB)", smem); return SNP_ERR_UNSUPPORTED; }
#define SNP_LAUNCH2(CTA_, PA_, HALF_, R2_)                                                                             \
    do {                                                                                                               \
        auto kern = k_step<T, SOC, OBS, HEADED, CTA_, PA_, HALF_, R2_>;                                                \
        if (smem > 48 * 1024) SNP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, block, smem, st>>>(a);                                                                            \
    } while (0)
#define SNP_LAUNCH(CTA_, PA_, HALF_) SNP_LAUNCH2(CTA_, PA_, HALF_, false)
    if (a.robot_mode >= 2) {  // robot driven by its own model, or a unicycle robot: separate instantiations so the common path keeps its registers
        if (per_agent) { set_error("robot_mode 2 with per-agent parameter rows is not supported"); return SNP_ERR_UNSUPPORTED; }
        if (!a.robot) { set_error("robot_mode 2 needs a robot array"); return SNP_ERR_INVALID; }
        if (cta) { if (half) SNP_LAUNCH2(true, false, true, true); else SNP_LAUNCH2(true, false, false, true); }
        else { if (half) SNP_LAUNCH2(false, false, true, true); else SNP_LAUNCH2(false, false, false, true); }
    } else if (cta) {
        if (per_agent) SNP_LAUNCH(true, true, false);
        else if (half) SNP_LAUNCH(true, false, true);
        else SNP_LAUNCH(true, false, false);
    } else if (per_agent) SNP_LAUNCH(false, true, false);
    else if (half) SNP_LAUNCH(false, false, true);
    else SNP_LAUNCH(false, false, false);
#undef SNP_LAUNCH2
#undef SNP_LAUNCH
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // namespace

template <typename T> int launch_step_small(const KArgs<T> &a, int type, cudaStream_t st) {
    switch (type) {
        case 0: return launch_one<T, 0, 0, 0>(a, st);
        case 1: return launch_one<T, 1, 1, 0>(a, st);
        case 2: return launch_one<T, 2, 0, 0>(a, st);
        case 3: return launch_one<T, 0, 0, 1>(a, st);
        case 4: return launch_one<T, 1, 1, 1>(a, st);
        case 5: return launch_one<T, 2, 0, 1>(a, st);
        case 6: return launch_one<T, 0, 0, 2>(a, st);
        case 7: return launch_one<T, 1, 1, 2>(a, st);
        case 8: return launch_one<T, 2, 0, 2>(a, st);
    }
    set_error("Type %d does not exist for this implementation", type);  // forces_parallel.py:211
    return SNP_ERR_INVALID;
}

template int launch_step_small<SNP_STEP_DTYPE>(const KArgs<SNP_STEP_DTYPE> &, int, cudaStream_t);

}  // namespace snp
