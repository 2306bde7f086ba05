// snp_laser.cu -- the laser range finder as a ray-vs-circle / ray-vs-segment intersection kernel.
//
// Reference: social_gym/src/sensors.py:53-69 get_laser_measurements (angles = linspace(yaw - range/2, yaw + range/2, samples),
// per ray the minimum over humans then over wall segments with strict '<'), :24-33 sphere_ray_intersect, :35-51
// segment_ray_intersect (one-sided: denominator <= 0 -> miss; 0 < t < 1, u > 0), src/robot_agent.py:81 (minus robot radius).
// The hit index is the first entity attaining the minimum in the reference's iteration order: humans 0..N-1, then the
// non-padding segment slots in (wall, slot) order offset by N; -1 when nothing is closer than max_distance.
//
// Two mappings:
//   k_laser_rays : one thread per ray, the env's circles and segments staged once in shared memory (small crowds);
//   k_laser_warp : one warp per ray, lanes stride over the entities and a warp-shuffle (value, index) min-reduction picks
//                  the winner, ties to the lowest index (large crowds).
#include "snp_kernels.cuh"

namespace snp {
namespace {

template <typename T> struct Circ { T sx, sy, cc; };       // s = sensor - centre, cc = s.s - r^2
// Exact-formula arithmetic for double (matches the oracle bit for bit given the same ray direction); plain for float.
template <typename T> struct Ops;
template <> struct Ops<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ void sincos(double a, double *s, double *c) { ::sincos(a, s, c); }
};
template <> struct Ops<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ void sincos(float a, float *s, float *c) { sincosf(a, s, c); }
};

// One wall segment as the ray loop wants it.  Everything that does not depend on the ray direction is evaluated once per env
// with the reference's own operations (sensors.py:36-47): a = x1-x2, b = y1-y2, e = x1-x3, f = y1-y3 (x3,y3 = sensor),
// un = -((x1-x2)*(y1-y3) - (y1-y2)*(x1-x3)) = the numerator of u.
template <typename T> struct RaySeg { T x1, y1, x2, y2, a, b, e, f, un; };

template <typename T> __device__ __forceinline__ RaySeg<T> make_rayseg(T x1, T y1, T x2, T y2, T x3, T y3) {
    using O = Ops<T>;
    RaySeg<T> s;
    s.x1 = x1; s.y1 = y1; s.x2 = x2; s.y2 = y2;
    s.a = O::sub(x1, x2); s.b = O::sub(y1, y2); s.e = O::sub(x1, x3); s.f = O::sub(y1, y3);
    s.un = -O::sub(O::mul(s.a, s.f), O::mul(s.b, s.e));
    return s;
}

template <typename T> __device__ __forceinline__ Circ<T> make_circ(T ox, T oy, T cx, T cy, T r) {
    using O = Ops<T>;
    Circ<T> c;
    c.sx = O::sub(ox, cx); c.sy = O::sub(oy, cy);
    c.cc = O::sub(O::fma(c.sy, c.sy, O::mul(c.sx, c.sx)), O::mul(r, r));  // np.dot(s,s) - r*r
    return c;
}

template <typename T> __device__ __forceinline__ T circle_hit(const Circ<T> &c, T dx, T dy, T maxd) {  // sensors.py:24-33
    using O = Ops<T>;
    const T b = O::fma(c.sy, dy, O::mul(c.sx, dx));  // np.dot(s, dir)
    T h = O::sub(O::mul(b, b), c.cc);
    if (h < T(0)) return maxd;
    h = Real<T>::sqrt_exact(h);
    const T t = O::sub(-b, h);
    if (t < T(0)) return maxd;
    return t < maxd ? t : maxd;
}

// sensors.py:35-51 for one (ray, segment): (c, d) = (y3 - y4, x3 - x4) are per-ray constants.  The two divisions of the
// reference (t and u) are only needed to LOCATE a hit: den > 0, 0 < t and u > 0 are decided exactly by the numerators' signs,
// and t < 1 implies nt < den, so the division (and the reference's own `t < 1` test on the rounded quotient) runs only for
// the few candidate hits.  Same hit/miss decisions and the same range bits as the division-first form.
template <typename T> __device__ __forceinline__ T segment_hit(const RaySeg<T> &s, T x3, T y3, T c, T d, T maxd) {
    using O = Ops<T>;
    const T den = O::sub(O::mul(s.a, c), O::mul(s.b, d));
    if (!(den > T(0))) return maxd;
    const T nt = O::sub(O::mul(s.e, c), O::mul(s.f, d));
    if (!(nt > T(0) && nt < den && s.un > T(0))) return maxd;
    const T t = O::div(nt, den);
    if (!(t < T(1))) return maxd;
    const T ix = O::add(s.x1, O::mul(t, O::sub(s.x2, s.x1))), iy = O::add(s.y1, O::mul(t, O::sub(s.y2, s.y1)));
    const T ex = O::sub(x3, ix), ey = O::sub(y3, iy);
    const T dist = Real<T>::sqrt_exact(O::fma(ey, ey, O::mul(ex, ex)));  // np.linalg.norm
    return dist < maxd ? dist : maxd;
}

template <typename T> __device__ __forceinline__ void ray_direction(T yaw, T range, int samples, int k, T &dx, T &dy) {
    // numpy.linspace(start, stop, samples): start + k*step with step = (stop-start)/(samples-1), last element = stop exactly
    using O = Ops<T>;
    const T half = O::div(range, T(2));
    const T start = O::sub(yaw, half), stop = O::add(yaw, half);
    const int div = samples - 1;
    T ang = start;
    if (div > 0) {
        const T step = O::div(O::sub(stop, start), T(div));
        ang = (k == div) ? stop : O::add(O::mul(T(k), step), start);
    }
    O::sincos(ang, &dy, &dx);
}

template <typename T> struct LArgs {
    int E, N, samples, W, S, walls_per_env;
    const T *px, *py, *radius, *walls, *pose;
    T range, maxd, robot_radius;
    T *ranges;
    int *hits;
    double sigma;                        // add_uncertainty (sensors.py:71-74); <= 0: off
    unsigned long long seed, scan;
};

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so ray (env, k) of scan s owns its random numbers whatever thread computes it.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// clip(N(m, sigma), 0, maxd): one Box-Muller normal from two 32-bit uniforms in (0, 1)
template <typename T> __device__ __forceinline__ T add_uncertainty(const LArgs<T> &a, T m, int e, int ray) {
    if (!(a.sigma > 0.0)) return m;
    unsigned r[4];
    philox4x32_10((unsigned)ray, (unsigned)e, (unsigned)a.scan, (unsigned)(a.scan >> 32), (unsigned)a.seed, (unsigned)(a.seed >> 32), r);
    const double u1 = ((double)r[0] + 0.5) * 2.3283064365386963e-10, u2 = ((double)r[1] + 0.5) * 2.3283064365386963e-10;
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    double v = fma(a.sigma, z, (double)m);
    v = v < (double)a.maxd ? v : (double)a.maxd;
    return (T)(v > 0.0 ? v : 0.0);
}

template <typename T> __global__ void __launch_bounds__(128) k_laser_rays(const LArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int e = blockIdx.y;
    const int nseg = a.W * a.S;
    Circ<T> *circ = reinterpret_cast<Circ<T> *>(smem_raw);
    RaySeg<T> *segs = reinterpret_cast<RaySeg<T> *>(smem_raw + ((sizeof(Circ<T>) * a.N + 15) & ~size_t(15)));
    const T ox = a.pose[e], oy = a.pose[(size_t)a.E + e], yaw = a.pose[2 * (size_t)a.E + e];
    for (int k = threadIdx.x; k < a.N; k += blockDim.x) {
        const size_t idx = (size_t)e * a.N + k;
        circ[k] = make_circ<T>(ox, oy, a.px[idx], a.py[idx], a.radius[idx]);
    }
    const T *w = a.walls + (a.walls_per_env ? (size_t)e * nseg * 4 : 0);
    for (int k = threadIdx.x; k < nseg; k += blockDim.x) segs[k] = make_rayseg<T>(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3], ox, oy);
    __syncthreads();
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= a.samples) return;
    T dx, dy;
    ray_direction<T>(yaw, a.range, a.samples, ray, dx, dy);
    const T rc_c = Ops<T>::sub(oy, Ops<T>::add(oy, dy)), rc_d = Ops<T>::sub(ox, Ops<T>::add(ox, dx));  // y3 - y4, x3 - x4 (sensors.py:42-44)
    T best = a.maxd;
    int hit = -1;
    for (int k = 0; k < a.N; ++k) {
        const T rc = circle_hit<T>(circ[k], dx, dy, a.maxd);
        if (rc < best) { best = rc; hit = k; }
    }
    int ord = a.N;
    for (int k = 0; k < nseg; ++k) {
        const RaySeg<T> &s = segs[k];
        if (s.x1 != s.x1) continue;  // NaN padding slot
        const T rc = segment_hit<T>(s, ox, oy, rc_c, rc_d, a.maxd);
        if (rc < best) { best = rc; hit = ord; }
        ++ord;
    }
    a.ranges[(size_t)e * a.samples + ray] = add_uncertainty<T>(a, best, e, ray) - a.robot_radius;
    if (a.hits) a.hits[(size_t)e * a.samples + ray] = hit;
}

// One warp per ray.  Segment ordinals need the count of non-padding slots before each slot; padding is a suffix of each
// wall's slots (motion_model_manager.py:270-275), so ordinal(w, s) = (valid slots of walls < w) + s, built in smem.
template <typename T> __global__ void __launch_bounds__(256) k_laser_warp(const LArgs<T> a) {
    extern __shared__ int wall_base[];  // [W] ordinal of each wall's first segment
    const int e = blockIdx.y;
    const int nseg = a.W * a.S;
    const T *w = a.walls + (a.walls_per_env ? (size_t)e * nseg * 4 : 0);
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < a.W; ++k) {
            wall_base[k] = acc;
            for (int s = 0; s < a.S; ++s) { const T x = w[4 * (k * a.S + s)]; acc += (x == x) ? 1 : 0; }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= a.samples) return;
    const T ox = a.pose[e], oy = a.pose[(size_t)a.E + e], yaw = a.pose[2 * (size_t)a.E + e];
    T dx, dy;
    ray_direction<T>(yaw, a.range, a.samples, ray, dx, dy);
    const T rc_c = Ops<T>::sub(oy, Ops<T>::add(oy, dy)), rc_d = Ops<T>::sub(ox, Ops<T>::add(ox, dx));
    T best = a.maxd;
    int hit = 0x7fffffff;
    for (int k = lane; k < a.N; k += 32) {
        const size_t idx = (size_t)e * a.N + k;
        const Circ<T> c = make_circ<T>(ox, oy, a.px[idx], a.py[idx], a.radius[idx]);
        const T rc = circle_hit<T>(c, dx, dy, a.maxd);
        if (rc < best) { best = rc; hit = k; }
    }
    for (int k = lane; k < nseg; k += 32) {
        const RaySeg<T> s = make_rayseg<T>(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3], ox, oy);
        if (s.x1 != s.x1) continue;
        const int wi = k / a.S;
        const T rc = segment_hit<T>(s, ox, oy, rc_c, rc_d, a.maxd);
        if (rc < best) { best = rc; hit = a.N + wall_base[wi] + (k - wi * a.S); }
    }
    // (value, index) min-reduction; ties -> lowest index == the reference's first strict-'<' winner
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oh = __shfl_xor_sync(0xffffffffu, hit, off);
        if (ob < best || (ob == best && oh < hit)) { best = ob; hit = oh; }
    }
    if (lane == 0) {
        a.ranges[(size_t)e * a.samples + ray] = add_uncertainty<T>(a, best, e, ray) - a.robot_radius;
        if (a.hits) a.hits[(size_t)e * a.samples + ray] = (hit == 0x7fffffff) ? -1 : hit;
    }
}

template <typename T> int run_laser(const snp_laser_args *g, cudaStream_t st) {
    LArgs<T> a;
    a.E = g->E; a.N = g->N; a.samples = g->samples; a.W = g->W; a.S = g->W > 0 ? g->S : 0; a.walls_per_env = g->walls_per_env;
    a.px = (const T *)g->px; a.py = (const T *)g->py; a.radius = (const T *)g->radius; a.walls = (const T *)g->walls; a.pose = (const T *)g->pose;
    a.range = (T)g->range; a.maxd = (T)g->max_distance; a.robot_radius = (T)g->robot_radius;
    a.ranges = (T *)g->ranges; a.hits = g->hits;
    a.sigma = g->uncertainty; a.seed = g->noise_seed; a.scan = g->noise_scan;
    const int entities = a.N + a.W * a.S;
    const size_t smem_rays = ((sizeof(Circ<T>) * a.N + 15) & ~size_t(15)) + sizeof(RaySeg<T>) * (size_t)(a.W * a.S);
    if (entities <= 512 && smem_rays <= 48 * 1024) {
        dim3 grid((a.samples + 127) / 128, a.E);
        k_laser_rays<T><<<grid, 128, smem_rays, st>>>(a);
    } else {
        dim3 grid((a.samples + 7) / 8, a.E);
        k_laser_warp<T><<<grid, 256, sizeof(int) * (a.W > 0 ? a.W : 1), st>>>(a);
    }
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // namespace
}  // namespace snp

using namespace snp;

extern "C" int snp_laser(const snp_laser_args *g, void *stream) {
    if (!g || !g->px || !g->py || !g->radius || !g->pose || !g->ranges) { set_error("snp_laser: null array"); return SNP_ERR_INVALID; }
    if (g->E <= 0 || g->N < 0 || g->samples <= 0) { set_error("snp_laser: E=%d N=%d samples=%d", g->E, g->N, g->samples); return SNP_ERR_INVALID; }
    if (g->E > 65535) { set_error("snp_laser: at most 65535 envs per call (got %d)", g->E); return SNP_ERR_INVALID; }
    if (g->max_distance > 10.0) { set_error("Maxium distance for laser is 10 meters"); return SNP_ERR_INVALID; }  // sensors.py:13
    if (g->W > 0 && !g->walls) { set_error("snp_laser: W=%d but walls is null", g->W); return SNP_ERR_INVALID; }
    if (g->dtype == SNP_F64) return run_laser<double>(g, (cudaStream_t)stream);
    if (g->dtype == SNP_F32) return run_laser<float>(g, (cudaStream_t)stream);
    set_error("snp_laser: bad dtype %d", g->dtype);
    return SNP_ERR_INVALID;
}
