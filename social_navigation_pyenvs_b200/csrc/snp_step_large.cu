// snp_step_large.cu -- one very large crowd: shared-memory tiled all-pairs (N-body style) social force + the same wall /
// desired / torque / Euler epilogue as the small-crowd kernel, one sub-step per launch.
//
// Every agent i of this rank's crowd accumulates the force of ALL M entities of the `others` view
// ([5][M] = x, y, vx, vy, r+safety in SoA, float4-coalesced tile loads) in ascending j -- the accumulation order of the
// reference's row-major pair loop (social_gym/src/forces.py:145-151), so the fp64 result matches the oracle to rounding.
// The new (x, y, vx, vy) of the own agents is written straight into `next_view` (double-buffered by the host), which is
// what the other ranks all-gather when the crowd is sharded by agent; own state is updated in place.
// Reference: same as snp_step_small.cu (motion_model_manager.py:354-373,424-459; forces.py).
#include "snp_kernels.cuh"

namespace snp {
namespace {

constexpr int kTile = 128;      // entities per shared-memory tile == threads per block
constexpr int kAgentsPerThread = 2;

template <typename T> struct LargeArgs {
    KArgs<T> k;
    const T *others;
    long long M, self_offset;
    T *next_view;
};

template <typename T, int SOC, int OBS, int HEADED>
__global__ void __launch_bounds__(kTile) k_large_step(const LargeArgs<T> la) {
    using R = Real<T>;
    const KArgs<T> &a = la.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nseg = a.W * a.S;
    double *exp_tbl_s = reinterpret_cast<double *>(smem_raw);
    Seg<T> *segs = reinterpret_cast<Seg<T> *>(smem_raw + 512);
    size_t off = 512 + ((sizeof(Seg<T>) * (size_t)nseg + 31) & ~size_t(31));
    Ent<T> *tile = reinterpret_cast<Ent<T> *>(smem_raw + off);
    T *tile_rs = reinterpret_cast<T *>(smem_raw + off + sizeof(Ent<T>) * kTile);
    int *seg_cnt = reinterpret_cast<int *>(smem_raw + off + sizeof(Ent<T>) * kTile + sizeof(T) * kTile);

    if (sizeof(T) == 8) exp_table_init(exp_tbl_s);
    for (int k = threadIdx.x; k < nseg; k += blockDim.x) {
        const T *w = a.walls + (size_t)k * 4;
        segs[k] = make_seg<T>(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.W; k += blockDim.x) {
        int c = 0;
        while (c < a.S && segs[(size_t)k * a.S + c].ax == segs[(size_t)k * a.S + c].ax) ++c;
        seg_cnt[k] = c;
    }

    const long long N = a.EN;
    const long long M = la.M;
    const Params<T> &P = a.P;
    Agent<T> me[kAgentsPerThread];
    long long idx[kAgentsPerThread];
    bool live[kAgentsPerThread];
    int gidx[kAgentsPerThread], gcnt[kAgentsPerThread];
    T fsx[kAgentsPerThread], fsy[kAgentsPerThread];
#pragma unroll
    for (int q = 0; q < kAgentsPerThread; ++q) {
        idx[q] = ((long long)blockIdx.x * kAgentsPerThread + q) * kTile + threadIdx.x;
        live[q] = idx[q] < N;
        const long long i = live[q] ? idx[q] : 0;
        Agent<T> &m = me[q];
        m.px = a.dyn[SNP_DYN_PX * N + i]; m.py = a.dyn[SNP_DYN_PY * N + i];
        m.vx = a.dyn[SNP_DYN_VX * N + i]; m.vy = a.dyn[SNP_DYN_VY * N + i];
        m.dfx = a.dyn[SNP_DYN_DFX * N + i]; m.dfy = a.dyn[SNP_DYN_DFY * N + i];
        if (HEADED) {
            m.th = a.dyn[SNP_DYN_TH * N + i]; m.bvx = a.dyn[SNP_DYN_BVX * N + i]; m.bvy = a.dyn[SNP_DYN_BVY * N + i];
            m.om = a.dyn[SNP_DYN_OM * N + i];
            R::sincos_(m.th, &m.sn, &m.cs);
            m.vx = np_mv(m.cs, -m.sn, m.bvx, m.bvy);
            m.vy = np_mv(m.sn, m.cs, m.bvx, m.bvy);
        } else { m.th = m.bvx = m.bvy = m.om = T(0); m.cs = T(1); m.sn = T(0); }
        m.r = a.stat[SNP_STAT_R * N + i]; m.m = a.stat[SNP_STAT_M * N + i]; m.vd = a.stat[SNP_STAT_VD * N + i];
        m.rs = m.r + a.stat[SNP_STAT_SAFETY * N + i];
        agent_static<T>(P, m);
        gidx[q] = a.goal_idx[i]; gcnt[q] = a.goal_cnt[i];
        m.gx = a.goals[((size_t)gidx[q] * 2 + 0) * N + i]; m.gy = a.goals[((size_t)gidx[q] * 2 + 1) * N + i];
        fsx[q] = T(0); fsy[q] = T(0);
    }

    // ---- tiled all-pairs ----
    const bool sym = a.symmetric != 0;
    for (long long j0 = 0; j0 < M; j0 += kTile) {
        __syncthreads();
        const long long j = j0 + threadIdx.x;
        if (j < M) {
            tile[threadIdx.x] = Ent<T>{la.others[j], la.others[M + j], la.others[2 * M + j], la.others[3 * M + j]};
            tile_rs[threadIdx.x] = la.others[4 * M + j];
        }
        __syncthreads();
        const int cnt = (int)min((long long)kTile, M - j0);
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const Ent<T> o = tile[t];
            const T rsj = tile_rs[t];
            const long long jj = j0 + t - la.self_offset;  // index of the entity in this crowd's numbering
#pragma unroll
            for (int q = 0; q < kAgentsPerThread; ++q) {
                Agent<T> &m = me[q];
                T fx, fy;  // the self pair (jj == idx[q]) contributes exactly zero by construction (tiny_ in pair_force)
                if (SOC == 2) {
                    const bool sw = sym && jj < idx[q];
                    pair_force<T, SOC>(P, exp_tbl_s, sw ? o.x : m.px, sw ? o.y : m.py, sw ? o.vx : m.vx, sw ? o.vy : m.vy, sw ? rsj : m.rs,
                                       sw ? m.px : o.x, sw ? m.py : o.y, sw ? m.vx : o.vx, sw ? m.vy : o.vy, sw ? m.rs : rsj, fx, fy);
                    fx = sw ? -fx : fx; fy = sw ? -fy : fy;
                } else {
                    pair_force<T, SOC>(P, exp_tbl_s, m.px, m.py, m.vx, m.vy, m.rs, o.x, o.y, o.vx, o.vy, rsj, fx, fy);
                }
                fsx[q] += fx; fsy[q] += fy;
            }
        }
    }

    // ---- epilogue per agent ----
#pragma unroll
    for (int q = 0; q < kAgentsPerThread; ++q) {
        if (!live[q]) continue;
        Agent<T> &m = me[q];
        const long long i = idx[q];
        const T dg = np_norm(m.gx - m.px, m.gy - m.py);
        if (a.numba ? (dg <= m.r) : (dg < m.r)) {
            gidx[q] = (gidx[q] + 1 >= gcnt[q]) ? 0 : gidx[q] + 1;
            m.gx = a.goals[((size_t)gidx[q] * 2 + 0) * N + i]; m.gy = a.goals[((size_t)gidx[q] * 2 + 1) * N + i];
        }
        T fox = T(0), foy = T(0);
        if (a.W > 0) obstacle_force<T, OBS>(P, exp_tbl_s, segs, seg_cnt, a.W, a.S, a.numba != 0, m.px, m.py, m.vx, m.vy, m.rs, fox, foy);
        desired_force<T>(P, m, a.numba != 0);
        integrate<T, HEADED>(P, m, fox, foy, fsx[q], fsy[q], a.dt);
        a.dyn[SNP_DYN_PX * N + i] = m.px; a.dyn[SNP_DYN_PY * N + i] = m.py;
        a.dyn[SNP_DYN_VX * N + i] = m.vx; a.dyn[SNP_DYN_VY * N + i] = m.vy;
        a.dyn[SNP_DYN_DFX * N + i] = m.dfx; a.dyn[SNP_DYN_DFY * N + i] = m.dfy;
        if (HEADED) {
            a.dyn[SNP_DYN_TH * N + i] = m.th; a.dyn[SNP_DYN_BVX * N + i] = m.bvx; a.dyn[SNP_DYN_BVY * N + i] = m.bvy;
            a.dyn[SNP_DYN_OM * N + i] = m.om;
        }
        a.goal_idx[i] = gidx[q];
        if (la.next_view) {
            const long long o = la.self_offset + i;
            la.next_view[o] = m.px; la.next_view[M + o] = m.py; la.next_view[2 * M + o] = m.vx; la.next_view[3 * M + o] = m.vy;
            la.next_view[4 * M + o] = m.rs;
        }
    }
}

// x, y, vx, vy, r+safety of every agent into a [5][stride] view (v = R(yaw) bv for headed models, mmm:448).
template <typename T> __global__ void k_large_publish(const T *dyn, const T *stat, long long N, int headed, T *view, long long stride, long long offset) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    T vx = dyn[SNP_DYN_VX * N + i], vy = dyn[SNP_DYN_VY * N + i];
    if (headed) {
        T s, c;
        Real<T>::sincos_(dyn[SNP_DYN_TH * N + i], &s, &c);
        const T bx = dyn[SNP_DYN_BVX * N + i], by = dyn[SNP_DYN_BVY * N + i];
        vx = np_mv(c, -s, bx, by); vy = np_mv(s, c, bx, by);
    }
    const long long o = offset + i;
    view[o] = dyn[SNP_DYN_PX * N + i]; view[stride + o] = dyn[SNP_DYN_PY * N + i];
    view[2 * stride + o] = vx; view[3 * stride + o] = vy;
    view[4 * stride + o] = stat[SNP_STAT_R * N + i] + stat[SNP_STAT_SAFETY * N + i];
}

template <typename T, int SOC, int OBS, int HEADED> int launch_large(const LargeArgs<T> &la, cudaStream_t st) {
    const long long N = la.k.EN;
    const int nseg = la.k.W * la.k.S;
    const size_t smem = 512 + ((sizeof(Seg<T>) * (size_t)nseg + 31) & ~size_t(31)) + sizeof(Ent<T>) * kTile + sizeof(T) * kTile + sizeof(int) * (la.k.W + 1) + 16;
    const long long per_block = (long long)kTile * kAgentsPerThread;
    const unsigned blocks = (unsigned)((N + per_block - 1) / per_block);
    auto kern = k_large_step<T, SOC, OBS, HEADED>;
    if (smem > 48 * 1024) SNP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kTile, smem, st>>>(la);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

template <typename T> int run_large(const snp_crowd *c, const snp_step_opts *o, const void *others, long long M, long long self_offset,
                                    void *next_view, cudaStream_t st) {
    LargeArgs<T> la;
    KArgs<T> &a = la.k;
    a.E = 1; a.N = 0; a.G = c->G; a.EN = (long long)c->E * c->N;
    a.dyn = (T *)c->dyn; a.stat = (const T *)c->stat; a.goals = (const T *)c->goals; a.goal_idx = c->goal_idx; a.goal_cnt = c->goal_cnt;
    a.agent_params = nullptr; a.P = make_params<T>(c->params); a.robot = nullptr;
    a.walls = (const T *)c->walls; a.W = c->W; a.S = c->W > 0 ? c->S : 0; a.walls_per_env = 0;
    a.consider_robot = 0; a.symmetric = o->symmetric; a.numba = o->numba_compat; a.n_substeps = 1; a.robot_mode = 0;
    a.dt = (T)o->dt; a.dt_d = o->dt; a.action = nullptr; a.pre_checks = a.post_checks = a.track_touch = 0;
    a.time_now = nullptr; a.flags = nullptr; a.checks = nullptr; a.epw = 1; a.full_pair_loop = 1;
    la.others = (const T *)others; la.M = M; la.self_offset = self_offset; la.next_view = (T *)next_view;
    switch (o->type) {
        case 0: return launch_large<T, 0, 0, 0>(la, st);
        case 1: return launch_large<T, 1, 1, 0>(la, st);
        case 2: return launch_large<T, 2, 0, 0>(la, st);
        case 3: return launch_large<T, 0, 0, 1>(la, st);
        case 4: return launch_large<T, 1, 1, 1>(la, st);
        case 5: return launch_large<T, 2, 0, 1>(la, st);
        case 6: return launch_large<T, 0, 0, 2>(la, st);
        case 7: return launch_large<T, 1, 1, 2>(la, st);
        case 8: return launch_large<T, 2, 0, 2>(la, st);
    }
    set_error("Type %d does not exist for this implementation", o->type);
    return SNP_ERR_INVALID;
}

}  // namespace
}  // namespace snp

using namespace snp;

extern "C" {

int snp_large_step(const snp_crowd *c, const snp_step_opts *o, const void *others, int64_t M, int64_t self_offset, void *next_view,
                   void *stream) {
    if (!c || !o || !others) { set_error("snp_large_step: null argument"); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->goals || !c->goal_idx || !c->goal_cnt) { set_error("snp_large_step: crowd arrays missing"); return SNP_ERR_INVALID; }
    if (c->agent_params) { set_error("snp_large_step: per-agent parameter rows are not supported"); return SNP_ERR_UNSUPPORTED; }
    if (c->walls_per_env) { set_error("snp_large_step: one wall set per crowd"); return SNP_ERR_INVALID; }
    const long long N = (long long)c->E * c->N;
    if (M < N || self_offset < 0 || self_offset + N > M + 1) { set_error("snp_large_step: M=%lld offset=%lld N=%lld", (long long)M, (long long)self_offset, N); return SNP_ERR_INVALID; }
    if (c->dtype == SNP_F64) return run_large<double>(c, o, others, M, self_offset, next_view, (cudaStream_t)stream);
    if (c->dtype == SNP_F32) return run_large<float>(c, o, others, M, self_offset, next_view, (cudaStream_t)stream);
    set_error("bad dtype %d", c->dtype);
    return SNP_ERR_INVALID;
}

int snp_large_publish(const snp_crowd *c, int32_t type, void *view, int64_t stride, int64_t offset, void *stream) {
    if (!c || !view || !c->dyn || !c->stat) { set_error("snp_large_publish: null argument"); return SNP_ERR_INVALID; }
    if (type < 0 || type > 8) { set_error("Type %d does not exist for this implementation", type); return SNP_ERR_INVALID; }
    const long long N = (long long)c->E * c->N;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    if (c->dtype == SNP_F64) k_large_publish<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double *)c->dyn, (const double *)c->stat, N, type >= 3, (double *)view, stride, offset);
    else k_large_publish<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)c->dyn, (const float *)c->stat, N, type >= 3, (float *)view, stride, offset);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // extern "C"
