// snp_layout.cu -- conversions between the reference's array-of-rows layout and the engine's structure-of-arrays.
//
// Reference rows (float64, social_gym/src/agent.py:256-258): [px,py,theta,vx,vy,bvx,bvy,omega,r,m,gx,gy,vd], robot (if
// visible) appended as row N of each env (motion_model_manager.py:259,359).  Goal lists are NaN padded [N][G][2]
// (motion_model_manager.py:262-267) and rotated in place by the reference; the engine keeps them fixed plus a head index.
// These kernels are pure HBM traffic: rows are staged through shared memory so that both the row side and the
// field side are accessed in contiguous, coalesced runs.
#include "snp_kernels.cuh"

namespace snp {
namespace {

constexpr int kRowsPerBlock = 128;
constexpr int kRow = 13;

template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock) k_unpack(const double *__restrict__ rows, int rows_per_env, const double *__restrict__ safety,
                                                           int E, int N, T *dyn, T *stat, T *robot) {
    __shared__ double tile[kRowsPerBlock * kRow];
    const long long total = (long long)E * rows_per_env;
    const long long r0 = (long long)blockIdx.x * kRowsPerBlock;
    const int nrows = (int)min((long long)kRowsPerBlock, total - r0);
    for (int k = threadIdx.x; k < nrows * kRow; k += kRowsPerBlock) tile[k] = rows[r0 * kRow + k];
    __syncthreads();
    if ((int)threadIdx.x >= nrows) return;
    const long long R = r0 + threadIdx.x;
    const long long e = R / rows_per_env;
    const int r = (int)(R - e * rows_per_env);
    const double *t = tile + threadIdx.x * kRow;
    const double saf = safety ? safety[R] : 0.0;
    const long long EN = (long long)E * N;
    if (r < N) {
        const long long a = e * N + r;
        dyn[SNP_DYN_PX * EN + a] = (T)t[0]; dyn[SNP_DYN_PY * EN + a] = (T)t[1]; dyn[SNP_DYN_TH * EN + a] = (T)t[2];
        dyn[SNP_DYN_VX * EN + a] = (T)t[3]; dyn[SNP_DYN_VY * EN + a] = (T)t[4];
        dyn[SNP_DYN_BVX * EN + a] = (T)t[5]; dyn[SNP_DYN_BVY * EN + a] = (T)t[6]; dyn[SNP_DYN_OM * EN + a] = (T)t[7];
        stat[SNP_STAT_R * EN + a] = (T)t[8]; stat[SNP_STAT_M * EN + a] = (T)t[9]; stat[SNP_STAT_VD * EN + a] = (T)t[12];
        stat[SNP_STAT_SAFETY * EN + a] = (T)saf;
    } else if (robot) {
        robot[SNP_ROBOT_PX * (long long)E + e] = (T)t[0]; robot[SNP_ROBOT_PY * (long long)E + e] = (T)t[1];
        robot[SNP_ROBOT_TH * (long long)E + e] = (T)t[2];
        robot[SNP_ROBOT_VX * (long long)E + e] = (T)t[3]; robot[SNP_ROBOT_VY * (long long)E + e] = (T)t[4];
        robot[SNP_ROBOT_R * (long long)E + e] = (T)t[8]; robot[SNP_ROBOT_SAFETY * (long long)E + e] = (T)saf;
        robot[SNP_ROBOT_GX * (long long)E + e] = (T)t[10]; robot[SNP_ROBOT_GY * (long long)E + e] = (T)t[11];
        robot[SNP_ROBOT_BVX * (long long)E + e] = (T)t[5]; robot[SNP_ROBOT_BVY * (long long)E + e] = (T)t[6];
        robot[SNP_ROBOT_OM * (long long)E + e] = (T)t[7]; robot[SNP_ROBOT_M * (long long)E + e] = (T)t[9];
        robot[SNP_ROBOT_VD * (long long)E + e] = (T)t[12];
    }
}

// Overwrites columns 0..7 and 10..11 of the human rows (what update_humans_parallel changes, fp:273-283,233-234);
// static columns and the robot row are left as the caller provided them.
template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock) k_pack(double *__restrict__ rows, int rows_per_env, int E, int N, int G, const T *dyn,
                                                         const T *goals, const int *goal_idx) {
    __shared__ double tile[kRowsPerBlock * kRow];
    const long long total = (long long)E * rows_per_env;
    const long long r0 = (long long)blockIdx.x * kRowsPerBlock;
    const int nrows = (int)min((long long)kRowsPerBlock, total - r0);
    for (int k = threadIdx.x; k < nrows * kRow; k += kRowsPerBlock) tile[k] = rows[r0 * kRow + k];
    __syncthreads();
    if ((int)threadIdx.x < nrows) {
        const long long R = r0 + threadIdx.x;
        const long long e = R / rows_per_env;
        const int r = (int)(R - e * rows_per_env);
        if (r < N) {
            const long long EN = (long long)E * N, a = e * N + r;
            double *t = tile + threadIdx.x * kRow;
            t[0] = (double)dyn[SNP_DYN_PX * EN + a]; t[1] = (double)dyn[SNP_DYN_PY * EN + a]; t[2] = (double)dyn[SNP_DYN_TH * EN + a];
            t[3] = (double)dyn[SNP_DYN_VX * EN + a]; t[4] = (double)dyn[SNP_DYN_VY * EN + a];
            t[5] = (double)dyn[SNP_DYN_BVX * EN + a]; t[6] = (double)dyn[SNP_DYN_BVY * EN + a]; t[7] = (double)dyn[SNP_DYN_OM * EN + a];
            const int gi = goal_idx[a];
            t[10] = (double)goals[((size_t)gi * 2 + 0) * EN + a]; t[11] = (double)goals[((size_t)gi * 2 + 1) * EN + a];
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nrows * kRow; k += kRowsPerBlock) rows[r0 * kRow + k] = tile[k];
}

// goals_rows [E][N][G][2] (NaN padded) -> goals SoA [G][2][EN], goal_cnt = slots before the first NaN, goal_idx = 0.
template <typename T>
__global__ void k_unpack_goals(const double *__restrict__ gr, long long EN, int G, T *goals, int *goal_idx, int *goal_cnt) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= EN) return;
    int cnt = 0;
    bool open = true;
    for (int k = 0; k < G; ++k) {
        const double x = gr[(a * G + k) * 2], y = gr[(a * G + k) * 2 + 1];
        if (open && x == x) ++cnt; else open = false;  // np.argwhere(np.isnan(goals[i]))[0][0]  (fp:227)
        goals[((size_t)k * 2 + 0) * EN + a] = (T)x; goals[((size_t)k * 2 + 1) * EN + a] = (T)y;
    }
    goal_idx[a] = 0;
    goal_cnt[a] = cnt > 0 ? cnt : 1;
}

// Writes the goal lists back in the reference's rotated form (list rotated left by goal_idx, fp:229-232) -- from the
// caller's own float64 rows so no precision is lost -- and resets nothing.
__global__ void k_rotate_goals(double *__restrict__ gr, long long EN, int G, const int *goal_idx, const int *goal_cnt, const int *prev_idx) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= EN) return;
    const int cnt = goal_cnt[a];
    const int sh = ((goal_idx[a] - (prev_idx ? prev_idx[a] : 0)) % cnt + cnt) % cnt;
    if (sh == 0) return;
    double tmp[2 * 16];
    if (cnt > 16) return;  // lists longer than 16 are rotated on the host side
    for (int k = 0; k < cnt; ++k) { tmp[2 * k] = gr[(a * G + k) * 2]; tmp[2 * k + 1] = gr[(a * G + k) * 2 + 1]; }
    for (int k = 0; k < cnt; ++k) { const int src = (k + sh) % cnt; gr[(a * G + k) * 2] = tmp[2 * src]; gr[(a * G + k) * 2 + 1] = tmp[2 * src + 1]; }
}

}  // namespace

template <typename T> int run_unpack(const snp_crowd *c, const double *rows, int rpe, const double *safety, cudaStream_t st) {
    const long long total = (long long)c->E * rpe;
    const unsigned blocks = (unsigned)((total + kRowsPerBlock - 1) / kRowsPerBlock);
    k_unpack<T><<<blocks, kRowsPerBlock, 0, st>>>(rows, rpe, safety, c->E, c->N, (T *)c->dyn, (T *)c->stat, rpe > c->N ? (T *)c->robot : nullptr);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

template <typename T> int run_pack(const snp_crowd *c, double *rows, int rpe, cudaStream_t st) {
    const long long total = (long long)c->E * rpe;
    const unsigned blocks = (unsigned)((total + kRowsPerBlock - 1) / kRowsPerBlock);
    k_pack<T><<<blocks, kRowsPerBlock, 0, st>>>(rows, rpe, c->E, c->N, c->G, (const T *)c->dyn, (const T *)c->goals, c->goal_idx);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

template <typename T> int run_unpack_goals(const snp_crowd *c, const double *gr, int *goal_cnt, cudaStream_t st) {
    const long long EN = (long long)c->E * c->N;
    k_unpack_goals<T><<<(unsigned)((EN + 255) / 256), 256, 0, st>>>(gr, EN, c->G, (T *)c->goals, c->goal_idx, goal_cnt);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

int run_rotate_goals(const snp_crowd *c, double *gr, cudaStream_t st) {
    const long long EN = (long long)c->E * c->N;
    k_rotate_goals<<<(unsigned)((EN + 255) / 256), 256, 0, st>>>(gr, EN, c->G, c->goal_idx, c->goal_cnt, nullptr);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // namespace snp

using namespace snp;

static int check_layout_args(const snp_crowd *c, const void *rows, int rpe) {
    if (!c || !rows) { set_error("null argument"); return SNP_ERR_INVALID; }
    if (rpe != c->N && rpe != c->N + 1) { set_error("rows_per_env must be N or N+1 (got %d, N=%d)", rpe, c->N); return SNP_ERR_INVALID; }
    if (rpe == c->N + 1 && !c->robot) { set_error("robot row present but crowd has no robot array"); return SNP_ERR_INVALID; }
    if (c->dtype != SNP_F32 && c->dtype != SNP_F64) { set_error("bad dtype %d", c->dtype); return SNP_ERR_INVALID; }
    return SNP_OK;
}

extern "C" {

int snp_unpack_states(const snp_crowd *c, const double *rows_dev, int32_t rpe, const double *safety_dev, void *stream) {
    int rc = check_layout_args(c, rows_dev, rpe);
    if (rc) return rc;
    return c->dtype == SNP_F64 ? run_unpack<double>(c, rows_dev, rpe, safety_dev, (cudaStream_t)stream)
                               : run_unpack<float>(c, rows_dev, rpe, safety_dev, (cudaStream_t)stream);
}

int snp_pack_states(const snp_crowd *c, double *rows_dev, int32_t rpe, void *stream) {
    int rc = check_layout_args(c, rows_dev, rpe);
    if (rc) return rc;
    return c->dtype == SNP_F64 ? run_pack<double>(c, rows_dev, rpe, (cudaStream_t)stream) : run_pack<float>(c, rows_dev, rpe, (cudaStream_t)stream);
}

int snp_unpack_goals(const snp_crowd *c, const double *goal_rows_dev, int32_t *goal_cnt_dev, void *stream) {
    if (!c || !goal_rows_dev || !goal_cnt_dev || !c->goals || !c->goal_idx) { set_error("null argument"); return SNP_ERR_INVALID; }
    return c->dtype == SNP_F64 ? run_unpack_goals<double>(c, goal_rows_dev, goal_cnt_dev, (cudaStream_t)stream)
                               : run_unpack_goals<float>(c, goal_rows_dev, goal_cnt_dev, (cudaStream_t)stream);
}

int snp_rotate_goal_rows(const snp_crowd *c, double *goal_rows_dev, void *stream) {
    if (!c || !goal_rows_dev || !c->goal_idx || !c->goal_cnt) { set_error("null argument"); return SNP_ERR_INVALID; }
    if (c->G > 16) { set_error("goal lists longer than 16 must be rotated by the host (G=%d)", c->G); return SNP_ERR_UNSUPPORTED; }
    return run_rotate_goals(c, goal_rows_dev, (cudaStream_t)stream);
}

}  // extern "C"
