// snp_api.cu -- extern "C" entry points of libsnp_b200.so (include/snp_b200.h): argument validation, translation of the
// plain-C descriptors into kernel argument blocks, error reporting and the launch counter.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include "snp_kernels.cuh"

namespace snp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {  // SM count of the CURRENT device (cached per device: one process may drive several)
    static std::atomic<int> sms[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cacheable = dev >= 0 && dev < 64;
    int n = cacheable ? sms[dev].load(std::memory_order_relaxed) : 0;
    if (n <= 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        if (cacheable) sms[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

template <typename T> static int build_args(const snp_crowd *c, const snp_step_opts *o, KArgs<T> &a) {
    if (!c || !o) { set_error("null descriptor"); return SNP_ERR_INVALID; }
    if (c->E <= 0 || c->N <= 0 || c->G <= 0) { set_error("E, N and G must be positive (E=%d N=%d G=%d)", c->E, c->N, c->G); return SNP_ERR_INVALID; }
    if (o->type < 0 || o->type > 8) { set_error("Type %d does not exist for this implementation", o->type); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->goals || !c->goal_idx || !c->goal_cnt) { set_error("dyn/stat/goals/goal_idx/goal_cnt must be device pointers"); return SNP_ERR_INVALID; }
    if (o->n_substeps < 0) { set_error("n_substeps must be >= 0"); return SNP_ERR_INVALID; }
    const bool need_robot = o->consider_robot || o->pre_checks || o->post_checks || o->track_touch || o->robot_mode;
    if (need_robot && !c->robot) { set_error("a robot array is required when consider_robot, robot_mode or any check is on"); return SNP_ERR_INVALID; }
    if (o->robot_mode < 0 || o->robot_mode > 3) { set_error("robot_mode must be 0, 1, 2 or 3"); return SNP_ERR_INVALID; }
    if ((o->robot_mode == 1 || o->robot_mode == 3 || o->pre_checks) && !o->action) { set_error("an action array is required for robot_mode 1 / 3 and pre_checks"); return SNP_ERR_INVALID; }
    if ((o->pre_checks || o->post_checks || o->track_touch) && !o->flags) { set_error("flags output required when checks are on"); return SNP_ERR_INVALID; }
    if (c->W < 0 || c->S < 0 || (c->W > 0 && (!c->walls || c->S == 0))) { set_error("walls: W=%d S=%d but no segment array", c->W, c->S); return SNP_ERR_INVALID; }
    if (c->agent_params && o->symmetric && !o->numba_compat) {
        set_error("per-agent parameter rows with the symmetric (all_equal_humans) path are not supported: pass a uniform row");
        return SNP_ERR_UNSUPPORTED;
    }
    a.E = c->E; a.N = c->N; a.G = c->G; a.EN = (long long)c->E * c->N;
    a.dyn = (T *)c->dyn; a.stat = (const T *)c->stat; a.goals = (const T *)c->goals;
    a.goal_idx = c->goal_idx; a.goal_cnt = c->goal_cnt;
    a.agent_params = (o->symmetric && o->numba_compat) ? nullptr : (const T *)c->agent_params;
    a.P = make_params<T>(c->params);
    a.robot = (T *)c->robot;
    a.walls = (const T *)c->walls; a.W = c->W; a.S = c->W > 0 ? c->S : 0; a.walls_per_env = c->walls_per_env;
    a.consider_robot = o->consider_robot; a.symmetric = o->symmetric; a.numba = o->numba_compat;
    a.n_substeps = o->n_substeps; a.robot_mode = o->robot_mode;
    a.dt = (T)o->dt; a.dt_d = o->dt;
    a.action = (const T *)o->action;
    a.pre_checks = o->pre_checks; a.post_checks = o->post_checks; a.track_touch = o->track_touch;
    for (int k = 0; k < 6; ++k) a.consts[k] = o->consts[k];
    a.time_now = o->time_now; a.flags = o->flags; a.checks = o->checks;
    a.robot_type = o->robot_type; a.RP = make_params<T>(o->robot_params);
    if (o->robot_mode == 2 && (o->robot_type < 0 || o->robot_type > 8)) { set_error("Type %d does not exist for this implementation", o->robot_type); return SNP_ERR_INVALID; }
    a.respawn = o->respawn; a.respawn_bounds[0] = o->respawn_bounds[0]; a.respawn_bounds[1] = o->respawn_bounds[1];
    if (o->respawn && c->N > 32) { set_error("parallel-traffic respawn is implemented for crowds of at most 32 humans per env"); return SNP_ERR_UNSUPPORTED; }
    a.epw = 1; a.gpb = 1;
    a.mapping = (o->reserved >> 2) & 3;  // bits 2-3 of `reserved`: thread mapping override (tests / tuning)
    a.full_pair_loop = o->reserved & 1;  // bit 0 of `reserved`: SNP_OPT_FULL_PAIR_LOOP
    a.dyn_out = (T *)o->dyn_out; a.goal_idx_out = o->goal_idx_out; a.respawn_envs = o->respawn_envs;
    a.robot_every = o->robot_mode == 2 ? o->robot_every : 0; a.robot_phase = o->robot_phase; a.robot_dt = (T)o->consts[5];
    if (a.robot_every < 0 || (a.robot_every > 1 && o->robot_phase < 0)) { set_error("robot_every / robot_phase must be >= 0"); return SNP_ERR_INVALID; }
    if (o->dyn_out && (o->respawn || o->robot_mode == 2)) { set_error("dyn_out (peek) cannot be combined with respawn or robot_mode 2"); return SNP_ERR_INVALID; }
    return SNP_OK;
}

// Device-side alias of a pinned host buffer (cudaHostAlloc / cudaHostRegister memory is mapped under unified addressing), or
// nullptr for pageable memory.
static void *device_alias(const void *host_ptr) {
    if (!host_ptr) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

template <typename T>
static int gym_step_zero_copy(const snp_crowd *crowd, const snp_step_opts *opts, void *obs_dev, int32_t *flags_dev, double *checks_dev, cudaStream_t st) {
    KArgs<T> a;
    const int rc = build_args<T>(crowd, opts, a);
    if (rc) return rc;
    a.obs_out = (T *)obs_dev; a.flags2 = flags_dev; a.checks2 = checks_dev;
    return launch_step_small<T>(a, opts->type, st);
}

}  // namespace snp

using namespace snp;

extern "C" {

int snp_abi_version(void) { return SNP_ABI_VERSION; }
const char *snp_last_error(void) { return g_err; }

int snp_device_info(int32_t *out4) {
    if (!out4) { set_error("null output"); return SNP_ERR_INVALID; }
    int dev = 0;
    SNP_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SNP_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    out4[0] = p.multiProcessorCount; out4[1] = p.major; out4[2] = p.minor; out4[3] = p.l2CacheSize;
    return SNP_OK;
}

int snp_step(const snp_crowd *crowd, const snp_step_opts *opts, void *stream) {
    if (!crowd) { set_error("null crowd"); return SNP_ERR_INVALID; }
    if (crowd->dtype == SNP_F64) {
        KArgs<double> a;
        int rc = build_args<double>(crowd, opts, a);
        if (rc) return rc;
        return launch_step_small<double>(a, opts->type, (cudaStream_t)stream);
    } else if (crowd->dtype == SNP_F32) {
        KArgs<float> a;
        int rc = build_args<float>(crowd, opts, a);
        if (rc) return rc;
        return launch_step_small<float>(a, opts->type, (cudaStream_t)stream);
    }
    set_error("dtype %d is neither SNP_F32 nor SNP_F64", crowd->dtype);
    return SNP_ERR_INVALID;
}

int snp_checks(const snp_crowd *crowd, const snp_step_opts *opts, void *stream) {
    if (!opts) { set_error("null opts"); return SNP_ERR_INVALID; }
    snp_step_opts o = *opts;
    o.n_substeps = 0;  // the fused kernel with an empty sub-step loop is exactly the checks (time is read, not advanced)
    o.robot_mode = 0;
    return snp_step(crowd, &o, stream);
}

int snp_gym_step_host(const snp_crowd *crowd, const snp_step_opts *opts, const void *action_host, void *obs_host, int32_t *flags_host,
                      double *checks_host, void *stream) {
    if (!crowd || !opts) { set_error("null descriptor"); return SNP_ERR_INVALID; }
    if (!opts->action) { set_error("snp_gym_step_host: opts->action must be the device action buffer [2][E]"); return SNP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t w = crowd->dtype == SNP_F64 ? 8 : 4;
    const size_t E = (size_t)crowd->E, EN = E * (size_t)crowd->N;
    if (action_host) SNP_CUDA_OK(cudaMemcpyAsync(const_cast<void *>(opts->action), action_host, 2 * E * w, cudaMemcpyHostToDevice, st));
    // Pinned result buffers: the kernel's store phase writes observation / flags / checks straight into them (posted PCIe writes,
    // overlapped with the CTAs still computing), so the only host-side cost after the launch is the synchronisation.  Pageable
    // buffers take the staged path below (D2H copies after the kernel).
    void *obs_dev = device_alias(obs_host);
    int32_t *flags_dev = (int32_t *)device_alias(flags_host);
    double *checks_dev = (double *)device_alias(checks_host);
    const bool zero_copy = opts->n_substeps > 0 && !opts->dyn_out && (!obs_host || obs_dev) && (!flags_host || flags_dev) && (!checks_host || checks_dev) &&
                           (obs_host || flags_host || checks_host) && (crowd->dtype == SNP_F64 || crowd->dtype == SNP_F32) && !(opts->reserved & SNP_OPT_STAGED_COPIES);
    if (zero_copy) {
        const int rc = crowd->dtype == SNP_F64 ? gym_step_zero_copy<double>(crowd, opts, obs_dev, flags_dev, checks_dev, st)
                                               : gym_step_zero_copy<float>(crowd, opts, obs_dev, flags_dev, checks_dev, st);
        if (rc) return rc;
        SNP_CUDA_OK(cudaStreamSynchronize(st));
        return SNP_OK;
    }
    const int rc = snp_step(crowd, opts, stream);
    if (rc) return rc;
    // px, py, vx, vy are the first four fields of dyn: one contiguous block
    if (obs_host) SNP_CUDA_OK(cudaMemcpyAsync(obs_host, crowd->dyn, 4 * EN * w, cudaMemcpyDeviceToHost, st));
    if (flags_host && opts->flags) SNP_CUDA_OK(cudaMemcpyAsync(flags_host, opts->flags, E * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (checks_host && opts->checks) SNP_CUDA_OK(cudaMemcpyAsync(checks_host, opts->checks, E * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SNP_CUDA_OK(cudaStreamSynchronize(st));
    return SNP_OK;
}

int64_t snp_launch_count(int32_t reset) {
    long long v = reset ? g_launches.exchange(0) : g_launches.load();
    return (int64_t)v;
}

}  // extern "C"
