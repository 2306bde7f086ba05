// snp_lookahead.cu -- the policy-side one-step lookahead producer: for every env and every action of the
// policy's action space, the swept collision / goal reward and the agent-centric ("rotated") joint state the value network
// consumes.  Reference: crowd_nav/policy/cadrl.py:42-83 compute_rotated_states_and_reward, :13-40
// transform_state_to_agent_centric (both Numba, float64), called once per decision from CADRL.predict (:235-262) after the
// peek get_next_human_observable_states (motion_model_manager.py:691-709, here snp_step with opts.dyn_out).
//
// Output volume dominates: [E][A][N][13|15] words (4096 x 81 x 25 x 13 doubles = 862 MB) against ~3 kB of input per env, so
// this is a pure HBM-WRITE kernel.  One CTA produces the rows of one (env, chunk of actions): a thread per (action, human)
// computes its 13 / 15 values into a shared-memory tile laid out exactly like the output, then the tile leaves as ONE bulk
// asynchronous copy (cp.async.bulk shared -> global, the TMA engine) -- no per-thread global stores, full 128-byte lines.
// The tile starts in shared memory at the same offset modulo 16 as its destination so that both sides of the bulk copy are
// 16-byte aligned; the (at most 15-byte) head and tail are written with scalar stores.
// Rewards are computed in double with formula-exact operations (snp_math.cuh) so that they are bit-identical to the
// reference for fp64 state; in fp32 mode from the widened fp32 state.
#include "snp_kernels.cuh"

namespace snp {
namespace {

constexpr int kLookThreads = 256;

struct LookArgs {
    int E, N, A, visible, headed;
    int chunk;   // actions per tile (one bulk copy)
    int aper;    // actions per CTA (a CTA walks its range tile by tile)
    long long EN;
    const void *dyn, *stat, *next, *robot;
    const double *actions;
    double dt;
    void *rotated;
    double *rewards;
    int bulk;  // 1: tiles leave through cp.async.bulk; 0: cooperative vector stores (A/B comparison)
};

// Per-action quantities shared by the humans of an env: next robot position, (cos, sin) of the rotation, distance to goal.
template <typename T> struct PerAction { double axd, ayd; T npx, npy, c, s, dg, ax, ay, pad; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <typename T> struct LookSmem {
    size_t pa, dist, tile, tile_stride, total;
    __host__ __device__ LookSmem(int N, int OW, int chunk, int aper) {
        size_t off = (sizeof(T) * 11 * (size_t)N + 15) & ~size_t(15);
        pa = off; off = (off + sizeof(PerAction<T>) * (size_t)aper + 15) & ~size_t(15);
        dist = off; off = (off + sizeof(double) * (size_t)aper * N + 15) & ~size_t(15);  // swept distances of the CTA's whole action range
        tile = off;
        tile_stride = ((size_t)chunk * N * OW * sizeof(T) + 16 + 15) & ~size_t(15);  // + 16: room for the alignment offset
        total = off + 2 * tile_stride;
    }
};

template <typename T>
__global__ void __launch_bounds__(kLookThreads, sizeof(T) == 4 ? 4 : 3) k_lookahead(const LookArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, OW = a.visible ? 15 : 13;
    const int e = blockIdx.y;
    const int abeg = blockIdx.x * a.aper;
    const int acnt = min(a.aper, a.A - abeg);
    const T *dyn = (const T *)a.dyn, *stat = (const T *)a.stat, *nxt = (const T *)a.next, *robot = (const T *)a.robot;
    const long long EN = a.EN, base = (long long)e * N;

    // shared memory: [cur: px py vx vy r | N each][next: x y vx vy th om | N each][per-action][dist x2][tile x2]
    const LookSmem<T> lay(N, OW, a.chunk, a.aper);
    T *cur = reinterpret_cast<T *>(smem_raw);
    T *nx = cur + 5 * N;
    PerAction<T> *pa = reinterpret_cast<PerAction<T> *>(smem_raw + lay.pa);
    double *dist_all = reinterpret_cast<double *>(smem_raw + lay.dist);

    // ---- stage the env's humans (coalesced SoA loads) and the robot ----
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        cur[0 * N + j] = dyn[SNP_DYN_PX * EN + base + j]; cur[1 * N + j] = dyn[SNP_DYN_PY * EN + base + j];
        cur[2 * N + j] = dyn[SNP_DYN_VX * EN + base + j]; cur[3 * N + j] = dyn[SNP_DYN_VY * EN + base + j];
        cur[4 * N + j] = stat[SNP_STAT_R * EN + base + j];
        nx[0 * N + j] = nxt[SNP_DYN_PX * EN + base + j]; nx[1 * N + j] = nxt[SNP_DYN_PY * EN + base + j];
        nx[2 * N + j] = nxt[SNP_DYN_VX * EN + base + j]; nx[3 * N + j] = nxt[SNP_DYN_VY * EN + base + j];
        // non-headed models do not integrate yaw / omega: the peek returns the current values (mmm:294-301)
        const T *ho = a.headed ? nxt : dyn;
        nx[4 * N + j] = ho[SNP_DYN_TH * EN + base + j]; nx[5 * N + j] = ho[SNP_DYN_OM * EN + base + j];
    }
    const long long E = a.E;
    const T rpx = robot[SNP_ROBOT_PX * E + e], rpy = robot[SNP_ROBOT_PY * E + e], rr = robot[SNP_ROBOT_R * E + e];
    const T rgx = robot[SNP_ROBOT_GX * E + e], rgy = robot[SNP_ROBOT_GY * E + e], rvd = robot[SNP_ROBOT_VD * E + e];
    // ---- per action, once per CTA: next robot position (cadrl.py:54), rotation (:22), distance to goal ----
    // (cos rot, sin rot) with rot = atan2(gy, gx) is the unit vector towards the goal: g / |g| (and (1, 0) for g = 0, atan2(0,0) = 0)
    // -- one division instead of an atan2 + sincos chain; agrees with the reference's libm round trip to ~1e-16.
    for (int k = threadIdx.x; k < acnt; k += blockDim.x) {
        PerAction<T> p;
        p.axd = a.actions[2 * (abeg + k)]; p.ayd = a.actions[2 * (abeg + k) + 1];
        p.ax = (T)p.axd; p.ay = (T)p.ayd; p.pad = T(0);
        if (sizeof(T) == 8) {  // no contraction: p + a * dt
            p.npx = (T)__dadd_rn((double)rpx, __dmul_rn(p.axd, a.dt)); p.npy = (T)__dadd_rn((double)rpy, __dmul_rn(p.ayd, a.dt));
            const double gx = __dsub_rn((double)rgx, (double)p.npx), gy = __dsub_rn((double)rgy, (double)p.npy);
            const double dg = xnorm_plain(gx, gy);
            p.dg = (T)dg;
            const double inv = dg > 0.0 ? 1.0 / dg : 0.0;
            p.c = (T)(dg > 0.0 ? gx * inv : 1.0); p.s = (T)(gy * inv);
        } else {
            const T dtT = (T)a.dt;
            p.npx = rpx + p.ax * dtT; p.npy = rpy + p.ay * dtT;
            const T gx = rgx - p.npx, gy = rgy - p.npy;
            p.dg = Real<T>::sqrt_exact(gx * gx + gy * gy);
            const T inv = p.dg > T(0) ? T(1) / p.dg : T(0);
            p.c = p.dg > T(0) ? gx * inv : T(1); p.s = gy * inv;
        }
        pa[k] = p;
    }
    __syncthreads();

    const size_t row_words = (size_t)N * OW;
    const int nchunks = (acnt + a.chunk - 1) / a.chunk;
    // A thread keeps the same (action slot k, human j) in every tile (a tile holds at most blockDim pairs unless N > blockDim):
    // the human's state and everything of the swept test that does not depend on the action stay in registers across tiles.
    const int k_own = threadIdx.x / N, j_own = threadIdx.x - k_own * N;
    const bool own = k_own < a.chunk;
    struct Human { T hr, nxx, nxy, nvx, nvy, th, om; double dx, dy, hvx, hvy, rsum; };
    auto load_human = [&](int j) {
        Human h;
        h.hr = cur[4 * N + j];
        h.nxx = nx[j]; h.nxy = nx[N + j]; h.nvx = nx[2 * N + j]; h.nvy = nx[3 * N + j]; h.th = nx[4 * N + j]; h.om = nx[5 * N + j];
        h.dx = __dsub_rn((double)cur[j], (double)rpx); h.dy = __dsub_rn((double)cur[N + j], (double)rpy);  // sim:962 difference
        h.hvx = (double)cur[2 * N + j]; h.hvy = (double)cur[3 * N + j];
        h.rsum = (double)h.hr;
        return h;
    };
    // one (action, human) pair: swept distance (social_nav_sim.py:962-976 via utils.py:22-36, formula-exact) and the rotated row
    auto do_pair = [&](const Human &h, const PerAction<T> &q, double *dist_out, T *o) {
        const double vx = __dsub_rn(h.hvx, q.axd), vy = __dsub_rn(h.hvy, q.ayd);
        const double ex = __dadd_rn(h.dx, __dmul_rn(vx, a.dt)), ey = __dadd_rn(h.dy, __dmul_rn(vy, a.dt));
        *dist_out = __dsub_rn(__dsub_rn(origin_to_segment(h.dx, h.dy, ex, ey), h.rsum), (double)rr);
        const T ddx = h.nxx - q.npx, ddy = h.nxy - q.npy;
        o[0] = q.dg; o[1] = rvd; o[2] = T(0); o[3] = rr;
        o[4] = q.ax * q.c + q.ay * q.s; o[5] = q.ay * q.c - q.ax * q.s;
        o[6] = ddx * q.c + ddy * q.s; o[7] = ddy * q.c - ddx * q.s;
        o[8] = h.nvx * q.c + h.nvy * q.s; o[9] = h.nvy * q.c - h.nvx * q.s;
        o[10] = h.hr;
        o[11] = Real<T>::sqrt_(ddx * ddx + ddy * ddy);  // < 1.5 ulp (snp_math.cuh); the rotated states carry the 1e-9 / 1e-4 tolerance
        o[12] = rr + h.hr;
        if (a.visible) { o[13] = h.th - T(0); o[14] = h.om; }
    };
    Human mine{};
    if (own) mine = load_human(j_own);

    for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * a.chunk;                    // first action of the tile, relative to abeg
        const int na = min(a.chunk, acnt - k0);
        const int pairs = na * N;
        T *gdst = (T *)a.rotated + ((size_t)e * a.A + abeg + k0) * row_words;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(gdst) & 15);
        // the tile sits in shared memory at the same offset modulo 16 as its destination: both sides of the copy 16-byte aligned
        T *tile = reinterpret_cast<T *>(smem_raw + lay.tile + (size_t)(c & 1) * lay.tile_stride + mis);
        double *dist = dist_all + (size_t)k0 * N;

        // ---- a thread per (action, human) ----
        if (own && k_own < na) do_pair(mine, pa[k0 + k_own], dist + threadIdx.x, tile + (size_t)threadIdx.x * OW);
        for (int p = threadIdx.x + blockDim.x; p < pairs; p += blockDim.x) {  // only when a tile holds more pairs than threads
            const int k = p / N, j = p - k * N;
            do_pair(load_human(j), pa[k0 + k], dist + p, tile + (size_t)p * OW);
        }
        const size_t total = (size_t)pairs * OW * sizeof(T);
        const size_t head = ((16 - mis) & 15) < total ? ((16 - mis) & 15) : total;
        const size_t body = (total - head) & ~size_t(15);
        const size_t tail = total - head - body;
        unsigned char *g8 = reinterpret_cast<unsigned char *>(gdst);
        const unsigned char *s8 = reinterpret_cast<const unsigned char *>(tile);
        if (a.bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy tile writes -> visible to the bulk copy
            // the copy of tile c-1 (other buffer) has had this whole compute phase to drain; the buffer is rewritten in iteration c+1
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();

        // ---- tile -> global ----
        if (a.bulk) {
            if (threadIdx.x == 0 && body) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g8 + head), "r"(smem_u32(s8 + head)), "r"((unsigned)body)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // head and tail (< 16 bytes each) in words of T
            const int hw = (int)(head / sizeof(T)), tw = (int)(tail / sizeof(T));
            if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + hw) gdst[threadIdx.x - 32] = tile[threadIdx.x - 32];
            if ((int)threadIdx.x >= 64 && (int)threadIdx.x < 64 + tw) {
                const size_t w = (head + body) / sizeof(T) + (threadIdx.x - 64);
                gdst[w] = tile[w];
            }
        } else {
            const int hw = (int)(head / sizeof(T));
            if ((int)threadIdx.x < hw) gdst[threadIdx.x] = tile[threadIdx.x];
            const uint4 *sv = reinterpret_cast<const uint4 *>(s8 + head);
            uint4 *gv = reinterpret_cast<uint4 *>(g8 + head);
            for (size_t v = threadIdx.x; v < body / 16; v += blockDim.x) gv[v] = sv[v];
            const int tw = (int)(tail / sizeof(T));
            if ((int)threadIdx.x < tw) { const size_t w = (head + body) / sizeof(T) + threadIdx.x; gdst[w] = tile[w]; }
            __syncthreads();  // single-phase fallback: the tile is reused two iterations later, the dist buffer too
        }

    }

    // ---- rewards (cadrl.py:56-72), once per CTA: a thread per action over the swept distances of all its humans.  (Doing this per
    //      tile made one warp late for every tile barrier: barrier stalls were the top stall reason of the first version.)  `break` in
    //      the reference only cuts its loop short: collision = some swept distance < 0, else dmin = the smallest one ----
    __syncthreads();
    for (int k = threadIdx.x; k < acnt; k += blockDim.x) {
        bool collision = false;
        double dmin = 9223372036854775807.0;  // np.iinfo(np.int64).max
#pragma unroll 5
        for (int j = 0; j < N; ++j) {
            const double d = dist_all[k * N + j];
            collision |= d < 0;
            dmin = (d >= 0 && d < dmin) ? d : dmin;
        }
        const PerAction<T> q = pa[k];
        const double npx = __dadd_rn((double)rpx, __dmul_rn(q.axd, a.dt)), npy = __dadd_rn((double)rpy, __dmul_rn(q.ayd, a.dt));
        const bool reached = xnorm_plain(__dsub_rn(npx, (double)rgx), __dsub_rn(npy, (double)rgy)) < (double)rr;
        double rew;
        if (collision) rew = -0.25;
        else if (reached) rew = 1.0;
        else if (dmin < 0.2) rew = __dmul_rn(__dmul_rn(__dsub_rn(dmin, 0.2), 0.5), a.dt);
        else rew = 0.0;
        a.rewards[(size_t)e * a.A + abeg + k] = rew;
    }
    if (a.bulk && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // tiles must outlive the copies' reads
}

template <typename T> int launch_lookahead(LookArgs a, cudaStream_t st) {
    const int N = a.N, OW = a.visible ? 15 : 13;
    // actions per tile: as many (action, human) pairs as threads, balanced over the tiles of a CTA; capped by shared memory
    int per = kLookThreads / N;
    if (per < 1) per = 1;
    const size_t row = (size_t)N * OW * sizeof(T);
    while (per > 1 && 2 * per * row > 64 * 1024) --per;
    // actions per CTA: whole envs when there are enough of them to fill the machine several times over, else split the action range
    const int sms = device_sm_count();
    int splits = 1;
    while (splits < a.A && (long long)a.E * splits < 8LL * 3 * sms) ++splits;
    for (;; ++splits) {  // ... and until the CTA's swept distances (aper x N doubles) and tiles fit in shared memory
        a.aper = (a.A + splits - 1) / splits;
        const int tiles = (a.aper + per - 1) / per;
        a.chunk = (a.aper + tiles - 1) / tiles;
        if (LookSmem<T>(N, OW, a.chunk, a.aper).total <= 200 * 1024 || a.aper == 1) break;
    }
    const LookSmem<T> lay(N, OW, a.chunk, a.aper);
    if (lay.total > 200 * 1024) { set_error("snp_lookahead: %d humans x %d values do not fit a shared-memory tile", N, OW); return SNP_ERR_UNSUPPORTED; }
    auto kern = k_lookahead<T>;
    if (lay.total > 48 * 1024) SNP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
    kern<<<dim3((unsigned)((a.A + a.aper - 1) / a.aper), (unsigned)a.E), kLookThreads, lay.total, st>>>(a);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // namespace
}  // namespace snp

using namespace snp;

namespace snp {
namespace {
// propagate_humans_state_with_constant_velocity_model (crowd_nav/policy/cadrl.py:92-105), the policies' query_env = False branch:
// x + vx dt, y + vy dt, theta + omega dt with the velocities carried over -- into a [SNP_DYN_FIELDS][E*N] buffer laid out like dyn,
// which is what snp_lookahead takes as `next`.  The reference's expression is a product followed by a sum (no contraction).
template <typename T> __global__ void k_constant_velocity(const T *dyn, T *next, long long EN, T dt) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= EN) return;
    const T vx = dyn[SNP_DYN_VX * EN + a], vy = dyn[SNP_DYN_VY * EN + a], om = dyn[SNP_DYN_OM * EN + a];
    if (sizeof(T) == 8) {
        next[SNP_DYN_PX * EN + a] = (T)__dadd_rn((double)dyn[SNP_DYN_PX * EN + a], __dmul_rn((double)vx, (double)dt));
        next[SNP_DYN_PY * EN + a] = (T)__dadd_rn((double)dyn[SNP_DYN_PY * EN + a], __dmul_rn((double)vy, (double)dt));
        next[SNP_DYN_TH * EN + a] = (T)__dadd_rn((double)dyn[SNP_DYN_TH * EN + a], __dmul_rn((double)om, (double)dt));
    } else {
        next[SNP_DYN_PX * EN + a] = (T)__fadd_rn((float)dyn[SNP_DYN_PX * EN + a], __fmul_rn((float)vx, (float)dt));
        next[SNP_DYN_PY * EN + a] = (T)__fadd_rn((float)dyn[SNP_DYN_PY * EN + a], __fmul_rn((float)vy, (float)dt));
        next[SNP_DYN_TH * EN + a] = (T)__fadd_rn((float)dyn[SNP_DYN_TH * EN + a], __fmul_rn((float)om, (float)dt));
    }
    next[SNP_DYN_VX * EN + a] = vx; next[SNP_DYN_VY * EN + a] = vy; next[SNP_DYN_OM * EN + a] = om;
    next[SNP_DYN_BVX * EN + a] = dyn[SNP_DYN_BVX * EN + a]; next[SNP_DYN_BVY * EN + a] = dyn[SNP_DYN_BVY * EN + a];
}
}  // namespace
}  // namespace snp

extern "C" int snp_constant_velocity(const snp_crowd *c, double dt, void *next, void *stream) {
    if (!c || !c->dyn || !next) { set_error("snp_constant_velocity: null argument"); return SNP_ERR_INVALID; }
    if (c->E <= 0 || c->N <= 0) { set_error("snp_constant_velocity: E and N must be positive"); return SNP_ERR_INVALID; }
    const long long EN = (long long)c->E * c->N;
    const unsigned blocks = (unsigned)((EN + 255) / 256);
    if (c->dtype == SNP_F64) snp::k_constant_velocity<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double *)c->dyn, (double *)next, EN, dt);
    else if (c->dtype == SNP_F32) snp::k_constant_velocity<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)c->dyn, (float *)next, EN, (float)dt);
    else { set_error("dtype %d is neither SNP_F32 nor SNP_F64", c->dtype); return SNP_ERR_INVALID; }
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

extern "C" int snp_lookahead(const snp_crowd *c, const snp_lookahead_args *g, void *stream) {
    if (!c || !g) { set_error("snp_lookahead: null descriptor"); return SNP_ERR_INVALID; }
    if (c->E <= 0 || c->N <= 0 || g->A <= 0) { set_error("snp_lookahead: E, N and A must be positive"); return SNP_ERR_INVALID; }
    if (c->E > 65535) { set_error("snp_lookahead: at most 65535 envs per call (got %d)", c->E); return SNP_ERR_UNSUPPORTED; }
    if (!c->dyn || !c->stat || !c->robot || !g->next || !g->actions || !g->rotated || !g->rewards) {
        set_error("snp_lookahead: dyn, stat, robot, next, actions, rotated and rewards must be device pointers");
        return SNP_ERR_INVALID;
    }
    if (g->type < 0 || g->type > 8) { set_error("Type %d does not exist for this implementation", g->type); return SNP_ERR_INVALID; }
    LookArgs a;
    a.E = c->E; a.N = c->N; a.A = g->A; a.visible = g->theta_and_omega_visible ? 1 : 0; a.headed = g->type >= 3; a.chunk = 1; a.aper = g->A;
    a.EN = (long long)c->E * c->N;
    a.dyn = c->dyn; a.stat = c->stat; a.next = g->next; a.robot = c->robot;
    a.actions = g->actions; a.dt = g->dt; a.rotated = g->rotated; a.rewards = g->rewards;
    a.bulk = (g->reserved & 1) ? 0 : 1;
    if (c->dtype == SNP_F64) return launch_lookahead<double>(a, (cudaStream_t)stream);
    if (c->dtype == SNP_F32) return launch_lookahead<float>(a, (cudaStream_t)stream);
    set_error("dtype %d is neither SNP_F32 nor SNP_F64", c->dtype);
    return SNP_ERR_INVALID;
}
