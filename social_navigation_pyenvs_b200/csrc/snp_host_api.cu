// snp_host_api.cu -- host-pointer entry points: the reference's operator signatures with a leading env axis.
//
//   snp_update_humans_parallel_host  <- update_humans_parallel(type, agents_state, goals, obstacles, agents_params, dt,
//                                        safety_space, all_params_equal, last_is_robot)   social_gym/src/forces_parallel.py:184
//   snp_laser_host                   <- LaserSensor.get_laser_measurements(humans, walls) social_gym/src/sensors.py:53
//
// Both copy the caller's float64 arrays to the device, run exactly the device-pointer kernels of the engine
// (unpack -> snp_step -> pack), copy the result back and synchronise.  Device workspaces are cached and only grow.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include "snp_kernels.cuh"

extern "C" int snp_unpack_goals(const snp_crowd *, const double *, int32_t *, void *);
extern "C" int snp_rotate_goal_rows(const snp_crowd *, double *, void *);

namespace snp {
namespace {

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t need) {
        if (need <= cap) return SNP_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = need + need / 4 + 256;
        SNP_CUDA_OK(cudaMalloc(&p, want));
        cap = want;
        return SNP_OK;
    }
};

struct Workspace {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    Buf rows, out_rows, goal_rows, safety, dyn, stat, goals, goal_idx, goal_cnt, agent_params, robot, walls, aos, pose, ranges, hits;
    int init() {
        if (!stream) SNP_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        return SNP_OK;
    }
};
// one workspace per device: buffers and the stream belong to the device that was current when they were created
Workspace g_ws_per_device[64];
Workspace &current_workspace() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_ws_per_device[(dev >= 0 && dev < 64) ? dev : 0];
}

// out[c*rows + r] = in[r*cols + c]
template <typename T> __global__ void k_aos_to_soa(const double *__restrict__ in, long long rows, int cols, T *__restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rows * cols) return;
    const long long r = k / cols;
    const int c = (int)(k - r * cols);
    out[(size_t)c * rows + r] = (T)in[k];
}
template <typename T> __global__ void k_soa_to_aos(const T *__restrict__ in, long long rows, int cols, double *__restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rows * cols) return;
    const long long r = k / cols;
    const int c = (int)(k - r * cols);
    out[k] = (double)in[(size_t)c * rows + r];
}
template <typename T> __global__ void k_convert(const double *__restrict__ in, long long n, T *__restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (T)in[k];
}

template <typename T>
int update_host(int type, int E, int N, int G, double *agents_state, double *goals, const double *obstacles, int W, int S,
                const double *agents_params, double dt, const double *safety_space, int all_params_equal, int last_is_robot,
                int numba_compat, int n_substeps, double *desired_force, double *out_state) {
    Workspace &ws = current_workspace();
    std::lock_guard<std::mutex> lock(ws.mu);
    int rc = ws.init();
    if (rc) return rc;
    cudaStream_t st = ws.stream;
    const int rows = N + (last_is_robot ? 1 : 0);
    const long long EN = (long long)E * N;
    const size_t n_rows = (size_t)E * rows * 13, n_goal = (size_t)EN * G * 2;
    const int dtype = sizeof(T) == 8 ? SNP_F64 : SNP_F32;

    if ((rc = ws.rows.ensure(n_rows * 8)) || (rc = ws.goal_rows.ensure(n_goal * 8)) || (rc = ws.safety.ensure((size_t)E * rows * 8)) ||
        (rc = ws.dyn.ensure(sizeof(T) * SNP_DYN_FIELDS * EN)) || (rc = ws.stat.ensure(sizeof(T) * SNP_STAT_FIELDS * EN)) ||
        (rc = ws.goals.ensure(sizeof(T) * n_goal)) || (rc = ws.goal_idx.ensure(4 * EN)) || (rc = ws.goal_cnt.ensure(4 * EN)) ||
        (rc = ws.robot.ensure(sizeof(T) * SNP_ROBOT_FIELDS * E)) || (rc = ws.aos.ensure((size_t)EN * 20 * 8)))
        return rc;

    snp_crowd c;
    memset(&c, 0, sizeof(c));
    c.E = E; c.N = N; c.G = G; c.dtype = dtype;
    c.dyn = ws.dyn.p; c.stat = ws.stat.p; c.goals = ws.goals.p; c.goal_idx = (int32_t *)ws.goal_idx.p; c.goal_cnt = (const int32_t *)ws.goal_cnt.p;
    c.robot = last_is_robot ? ws.robot.p : nullptr;

    // parameters: one uniform row when all rows are identical (always the case in the reference, agent.py:79-243)
    bool uniform = true;
    for (long long k = 1; k < EN && uniform; ++k) uniform = memcmp(agents_params, agents_params + k * 20, 20 * sizeof(double)) == 0;
    memcpy(c.params, agents_params, 20 * sizeof(double));
    if (!uniform) {
        if ((rc = ws.agent_params.ensure(sizeof(T) * 20 * EN))) return rc;
        SNP_CUDA_OK(cudaMemcpyAsync(ws.aos.p, agents_params, (size_t)EN * 20 * 8, cudaMemcpyHostToDevice, st));
        k_aos_to_soa<T><<<(unsigned)((EN * 20 + 255) / 256), 256, 0, st>>>((const double *)ws.aos.p, EN, 20, (T *)ws.agent_params.p);
        count_launch();
        c.agent_params = ws.agent_params.p;
    }
    // walls [W][S][2][2] -> [W*S][4] in the kernel's precision
    std::vector<T> wall_host;
    if (obstacles && W > 0 && S > 0) {
        wall_host.resize((size_t)W * S * 4);
        for (size_t k = 0; k < wall_host.size(); ++k) wall_host[k] = (T)obstacles[k];
        if ((rc = ws.walls.ensure(sizeof(T) * wall_host.size()))) return rc;
        SNP_CUDA_OK(cudaMemcpyAsync(ws.walls.p, wall_host.data(), sizeof(T) * wall_host.size(), cudaMemcpyHostToDevice, st));
        c.walls = ws.walls.p; c.W = W; c.S = S;
    }
    SNP_CUDA_OK(cudaMemcpyAsync(ws.rows.p, agents_state, n_rows * 8, cudaMemcpyHostToDevice, st));
    SNP_CUDA_OK(cudaMemcpyAsync(ws.goal_rows.p, goals, n_goal * 8, cudaMemcpyHostToDevice, st));
    SNP_CUDA_OK(cudaMemcpyAsync(ws.safety.p, safety_space, (size_t)E * rows * 8, cudaMemcpyHostToDevice, st));
    if ((rc = snp_unpack_states(&c, (const double *)ws.rows.p, rows, (const double *)ws.safety.p, st))) return rc;
    if ((rc = snp_unpack_goals(&c, (const double *)ws.goal_rows.p, (int32_t *)ws.goal_cnt.p, st))) return rc;
    T *df = (T *)ws.dyn.p + (size_t)SNP_DYN_DFX * EN;
    if (desired_force) {
        SNP_CUDA_OK(cudaMemcpyAsync(ws.aos.p, desired_force, (size_t)EN * 2 * 8, cudaMemcpyHostToDevice, st));
        k_aos_to_soa<T><<<(unsigned)((EN * 2 + 255) / 256), 256, 0, st>>>((const double *)ws.aos.p, EN, 2, df);
        count_launch();
    } else {
        SNP_CUDA_OK(cudaMemsetAsync(df, 0, sizeof(T) * 2 * EN, st));
    }

    snp_step_opts o;
    memset(&o, 0, sizeof(o));
    o.type = type; o.consider_robot = last_is_robot; o.symmetric = all_params_equal; o.numba_compat = numba_compat;
    o.n_substeps = n_substeps; o.dt = dt;
    if ((rc = snp_step(&c, &o, st))) return rc;

    // results: updated rows (np.copy of the input with the changed columns), rotated goal lists, carried desired force
    if ((rc = snp_pack_states(&c, (double *)ws.rows.p, rows, st))) return rc;
    SNP_CUDA_OK(cudaMemcpyAsync(out_state, ws.rows.p, n_rows * 8, cudaMemcpyDeviceToHost, st));
    if (G <= 16) {
        if ((rc = snp_rotate_goal_rows(&c, (double *)ws.goal_rows.p, st))) return rc;
        SNP_CUDA_OK(cudaMemcpyAsync(goals, ws.goal_rows.p, n_goal * 8, cudaMemcpyDeviceToHost, st));
    }
    std::vector<int> idx_host;
    if (G > 16) {
        idx_host.resize(EN);
        SNP_CUDA_OK(cudaMemcpyAsync(idx_host.data(), ws.goal_idx.p, 4 * EN, cudaMemcpyDeviceToHost, st));
    }
    if (desired_force) {
        k_soa_to_aos<T><<<(unsigned)((EN * 2 + 255) / 256), 256, 0, st>>>(df, EN, 2, (double *)ws.aos.p);
        count_launch();
        SNP_CUDA_OK(cudaMemcpyAsync(desired_force, ws.aos.p, (size_t)EN * 2 * 8, cudaMemcpyDeviceToHost, st));
    }
    SNP_CUDA_OK(cudaStreamSynchronize(st));
    // in-place side effects of the reference on its INPUT array (fp:254-256): linear velocity of headed agents
    // becomes R(theta) bv (old theta, old bv).  After the synchronisation: with pinned host memory the upload of agents_state is
    // asynchronous, so the rows must not change while it may still be in flight.  (out_state == agents_state is rejected above.)
    if (type >= 3) {
        for (int e = 0; e < E; ++e)
            for (int i = 0; i < N; ++i) {
                double *r = agents_state + ((size_t)e * rows + i) * 13;
                const double cs = std::cos(r[2]), sn = std::sin(r[2]);
                r[3] = cs * r[5] + -sn * r[6];
                r[4] = sn * r[5] + cs * r[6];
            }
    }
    if (G > 16) {  // long goal lists: rotate on the host
        std::vector<double> tmp(2 * (size_t)G);
        for (long long a = 0; a < EN; ++a) {
            double *g = goals + (size_t)a * G * 2;
            int cnt = 0;
            while (cnt < G && g[2 * cnt] == g[2 * cnt]) ++cnt;
            const int sh = cnt ? idx_host[a] % cnt : 0;
            if (!sh) continue;
            memcpy(tmp.data(), g, sizeof(double) * 2 * cnt);
            for (int k = 0; k < cnt; ++k) { g[2 * k] = tmp[2 * ((k + sh) % cnt)]; g[2 * k + 1] = tmp[2 * ((k + sh) % cnt) + 1]; }
        }
    }
    // goal columns of the input rows are refreshed too (fp:233)
    for (int e = 0; e < E; ++e)
        for (int i = 0; i < N; ++i) {
            const size_t k = ((size_t)e * rows + i) * 13;
            agents_state[k + 10] = out_state[k + 10];
            agents_state[k + 11] = out_state[k + 11];
        }
    return SNP_OK;
}

template <typename T>
int laser_host(int E, int N, const double *humans, const double *walls, int W, int S, const double *pose, double range, int samples,
               double max_distance, double robot_radius, double *ranges, int32_t *hits) {
    Workspace &ws = current_workspace();
    std::lock_guard<std::mutex> lock(ws.mu);
    int rc = ws.init();
    if (rc) return rc;
    cudaStream_t st = ws.stream;
    const long long EN = (long long)E * N;
    const size_t n_out = (size_t)E * samples;
    if ((rc = ws.aos.ensure((size_t)EN * 3 * 8 + (size_t)E * 3 * 8 + 64)) || (rc = ws.dyn.ensure(sizeof(T) * 3 * (EN > 0 ? EN : 1))) ||
        (rc = ws.pose.ensure(sizeof(T) * 3 * E)) || (rc = ws.ranges.ensure(n_out * 8 + sizeof(T) * n_out)) || (rc = ws.hits.ensure(n_out * 4)))
        return rc;
    double *d_h = (double *)ws.aos.p, *d_p = d_h + (size_t)EN * 3;
    if (EN > 0) {
        SNP_CUDA_OK(cudaMemcpyAsync(d_h, humans, (size_t)EN * 3 * 8, cudaMemcpyHostToDevice, st));
        k_aos_to_soa<T><<<(unsigned)((EN * 3 + 255) / 256), 256, 0, st>>>(d_h, EN, 3, (T *)ws.dyn.p);
        count_launch();
    }
    SNP_CUDA_OK(cudaMemcpyAsync(d_p, pose, (size_t)E * 3 * 8, cudaMemcpyHostToDevice, st));
    k_aos_to_soa<T><<<(unsigned)((E * 3 + 255) / 256), 256, 0, st>>>(d_p, E, 3, (T *)ws.pose.p);
    count_launch();
    snp_laser_args g;
    memset(&g, 0, sizeof(g));
    std::vector<T> wall_host;
    if (walls && W > 0 && S > 0) {
        wall_host.resize((size_t)W * S * 4);
        for (size_t k = 0; k < wall_host.size(); ++k) wall_host[k] = (T)walls[k];
        if ((rc = ws.walls.ensure(sizeof(T) * wall_host.size()))) return rc;
        SNP_CUDA_OK(cudaMemcpyAsync(ws.walls.p, wall_host.data(), sizeof(T) * wall_host.size(), cudaMemcpyHostToDevice, st));
        g.walls = ws.walls.p; g.W = W; g.S = S;
    }
    T *d_rng_t = (T *)((char *)ws.ranges.p + n_out * 8);
    g.E = E; g.N = N; g.dtype = sizeof(T) == 8 ? SNP_F64 : SNP_F32; g.samples = samples;
    g.px = ws.dyn.p; g.py = (T *)ws.dyn.p + EN; g.radius = (T *)ws.dyn.p + 2 * EN;
    g.pose = ws.pose.p; g.range = range; g.max_distance = max_distance; g.robot_radius = robot_radius;
    g.ranges = sizeof(T) == 8 ? ws.ranges.p : (void *)d_rng_t; g.hits = hits ? (int32_t *)ws.hits.p : nullptr;
    if ((rc = snp_laser(&g, st))) return rc;
    if (sizeof(T) != 8) {
        k_soa_to_aos<T><<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(d_rng_t, (long long)n_out, 1, (double *)ws.ranges.p);
        count_launch();
    }
    SNP_CUDA_OK(cudaMemcpyAsync(ranges, ws.ranges.p, n_out * 8, cudaMemcpyDeviceToHost, st));
    if (hits) SNP_CUDA_OK(cudaMemcpyAsync(hits, ws.hits.p, n_out * 4, cudaMemcpyDeviceToHost, st));
    SNP_CUDA_OK(cudaStreamSynchronize(st));
    return SNP_OK;
}

// ---- pipe-peak microbenchmarks: the measured denominators of the compute roofline ----
template <typename T> __global__ void __launch_bounds__(256) k_fma_peak(T *out, int iters, T a, T b) {
    T x0 = T(threadIdx.x), x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma_<T>(x0, a, b); x1 = fma_<T>(x1, a, b); x2 = fma_<T>(x2, a, b); x3 = fma_<T>(x3, a, b);
        x4 = fma_<T>(x4, a, b); x5 = fma_<T>(x5, a, b); x6 = fma_<T>(x6, a, b); x7 = fma_<T>(x7, a, b);
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == T(-1)) out[0] = x0;
}
__global__ void __launch_bounds__(256) k_mufu_peak(float *out, int iters) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    for (int i = 0; i < iters; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
    }
    if (x0 + x1 + x2 + x3 == -1.0f) out[0] = x0;
}

__global__ void k_debug_exp(const double *x, double *y, int n) {
    __shared__ __align__(16) double tbl[kExpN];
    exp_table_init(tbl);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = exp_tbl(x[i], tbl);
}

__global__ void k_debug_math(int kind, const double *x, const double *y, double *out, int n) {
    __shared__ __align__(16) double tbl[kExpN];
    exp_table_init(tbl);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == 0) out[i] = exp2_scaled(x[i], tbl);
    else if (kind == 1) out[i] = atan2_poly(y[i], x[i]);
    else if (kind == 2) sincos_bounded(x[i], &out[i], &out[n + i]);
    else if (kind == 3) out[i] = Real<double>::rsqrt_(x[i]);
    else out[i] = Real<double>::clamp01_(x[i]);
}

}  // namespace
}  // namespace snp

using namespace snp;

extern "C" {

int snp_debug_exp(const double *x_dev, double *y_dev, int32_t n, void *stream) {
    if (!x_dev || !y_dev || n <= 0) { set_error("snp_debug_exp: bad argument"); return SNP_ERR_INVALID; }
    SNP_CUDA_OK(ensure_exp_table());
    k_debug_exp<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, n);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

int snp_debug_math(int32_t kind, const double *x_dev, const double *y_dev, double *out_dev, int32_t n, void *stream) {
    if (!x_dev || !out_dev || n <= 0 || kind < 0 || kind > 4 || (kind == 1 && !y_dev)) { set_error("snp_debug_math: bad argument"); return SNP_ERR_INVALID; }
    SNP_CUDA_OK(ensure_exp_table());
    k_debug_math<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(kind, x_dev, y_dev, out_dev, n);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

int snp_update_humans_parallel_host(int32_t type, int32_t E, int32_t N, int32_t G, double *agents_state, double *goals,
                                    const double *obstacles, int32_t W, int32_t S, const double *agents_params, double dt,
                                    const double *safety_space, int32_t all_params_equal, int32_t last_is_robot, int32_t numba_compat,
                                    int32_t dtype, int32_t n_substeps, double *desired_force, double *out_state) {
    if (type < 0 || type > 8) { set_error("Type %d does not exist for this implementation", type); return SNP_ERR_INVALID; }
    if (E <= 0 || N <= 0 || G <= 0) { set_error("E, N, G must be positive"); return SNP_ERR_INVALID; }
    if (!agents_state || !goals || !agents_params || !safety_space || !out_state) { set_error("null array"); return SNP_ERR_INVALID; }
    if (n_substeps < 1) { set_error("n_substeps must be >= 1"); return SNP_ERR_INVALID; }
    if (out_state == agents_state) { set_error("out_state must not alias agents_state (the reference returns a new array, forces_parallel.py:214)"); return SNP_ERR_INVALID; }
    if (dtype == SNP_F64)
        return update_host<double>(type, E, N, G, agents_state, goals, obstacles, W, S, agents_params, dt, safety_space, all_params_equal,
                                   last_is_robot, numba_compat, n_substeps, desired_force, out_state);
    if (dtype == SNP_F32)
        return update_host<float>(type, E, N, G, agents_state, goals, obstacles, W, S, agents_params, dt, safety_space, all_params_equal,
                                  last_is_robot, numba_compat, n_substeps, desired_force, out_state);
    set_error("bad dtype %d", dtype);
    return SNP_ERR_INVALID;
}

int snp_laser_host(int32_t E, int32_t N, const double *humans, const double *walls, int32_t W, int32_t S, const double *pose, double range,
                   int32_t samples, double max_distance, double robot_radius, int32_t dtype, double *ranges, int32_t *hits) {
    if (E <= 0 || N < 0 || samples <= 0 || !pose || !ranges || (N > 0 && !humans)) { set_error("snp_laser_host: bad argument"); return SNP_ERR_INVALID; }
    if (dtype == SNP_F64) return laser_host<double>(E, N, humans, walls, W, S, pose, range, samples, max_distance, robot_radius, ranges, hits);
    if (dtype == SNP_F32) return laser_host<float>(E, N, humans, walls, W, S, pose, range, samples, max_distance, robot_radius, ranges, hits);
    set_error("bad dtype %d", dtype);
    return SNP_ERR_INVALID;
}

int snp_measure_pipe_peak(int32_t kind, double *out) {
    if (!out || kind < 0 || kind > 2) { set_error("snp_measure_pipe_peak: kind must be 0 (fp32 FMA), 1 (fp64 FMA) or 2 (MUFU.EX2)"); return SNP_ERR_INVALID; }
    const int sms = device_sm_count();
    const int blocks = sms * 8, threads = 256, iters = kind == 1 ? 1 << 14 : (kind == 0 ? 1 << 15 : 1 << 13);  // ~2 ms per launch each
    void *buf = nullptr;
    SNP_CUDA_OK(cudaMalloc(&buf, 64));
    cudaEvent_t e0, e1;
    SNP_CUDA_OK(cudaEventCreate(&e0));
    SNP_CUDA_OK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        SNP_CUDA_OK(cudaEventRecord(e0, 0));
        if (kind == 0) k_fma_peak<float><<<blocks, threads>>>((float *)buf, iters, 1.0000001f, 1e-7f);
        else if (kind == 1) k_fma_peak<double><<<blocks, threads>>>((double *)buf, iters, 1.0000001, 1e-7);
        else k_mufu_peak<<<blocks, threads>>>((float *)buf, iters);
        SNP_CUDA_OK(cudaEventRecord(e1, 0));
        SNP_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        SNP_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    const double ops = (double)blocks * threads * iters * (kind == 2 ? 4.0 : 16.0);
    *out = ops / (best_ms * 1e-3) / (kind == 2 ? 1e9 : 1e12);  // Gop/s for MUFU, TFLOP/s for FMA
    return SNP_OK;
}

}  // extern "C"
