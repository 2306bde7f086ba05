// snp_physics.cuh -- the SFM / HSFM force laws and the explicit-Euler update, written once and shared by the
// small-crowd fused kernel, the CTA-per-env kernel and the large-crowd tiled kernel.
//
// Reference (paths under social_gym/): src/forces.py:9-16 desired, :27-53 obstacle (Helbing / Guo), :63-128 pair laws
// (Helbing / Guo / Moussaid), :279-290 torque (Farina / "new"); src/motion_model_manager.py:424-435 global force,
// :72-85 Euler, :52-55 speed clip; src/obstacle.py:53-66 closest point on a polygon's segments.
#pragma once
#include "snp_math.cuh"

namespace snp {

// Parameter row of agent.py:269 in the form the kernels consume.  Every exponential of the force laws is A exp(x / B): the
// amplitude and the decay length are folded into the exponent,  A exp(x / B) = E(fma(x, kB, lA))  with  E(t) = exp(t / s),
// s = Real<T>::escale()  (fp64: the exp table's L/ln2; fp32: log2 e, so E is a bare ex2.approx),  kB = s / B,  lA = s ln A.
// Amplitudes are >= 0 in every parameter set of the reference (agent.py:269-312); A <= 0 is treated as 0 (lA = -1e7).
// The inertia of the torque law cancels: omega' = torque / I with k_theta = I k_lambda |f|, k_omega = I (1 + alpha)
// sqrt(k_lambda |f| / alpha)  (forces.py:279-290)  =>  torque / I = -k_lambda |f| dtheta - c_omega sqrt(|f|) omega.
template <typename T> struct Params {
    T inv_relax, lAi, kBi, lAw, kBw, lCi, kDi, lCw, kDw, lEi, k1, k2, lambda, gamma, ns, ns1, ko, kd, k_lambda, c_omega;
};

template <typename T> __host__ __device__ inline Params<T> make_params(const double *p) {
    const double s = (double)Real<T>::escale();
    auto lg = [s](double a) { return a > 0.0 ? s * log(a) : -1.0e7; };
    auto kk = [s](double b) { return b != 0.0 ? s / b : 0.0; };
    Params<T> q;
    q.inv_relax = T(1.0 / p[0]);
    q.lAi = T(lg(p[1])); q.lAw = T(lg(p[2]));
    q.kBi = T(kk(p[3])); q.kBw = T(kk(p[4]));
    q.lCi = T(lg(p[5])); q.lCw = T(lg(p[6]));
    q.kDi = T(kk(p[7])); q.kDw = T(kk(p[8]));
    q.lEi = T(lg(p[9])); q.k1 = T(p[10]); q.k2 = T(p[11]); q.lambda = T(p[12]); q.gamma = T(p[13]); q.ns = T(p[14]); q.ns1 = T(p[15]);
    q.ko = T(p[16]); q.kd = T(p[17]);
    q.k_lambda = T(p[19]);
    q.c_omega = T(p[18] != 0.0 ? (1.0 + p[18]) * sqrt(p[19] / p[18]) : 0.0);
    return q;
}

// Force exerted on agent 1 by agent 2 (forces.py:63-128).  rs = radius + safety_space.
// SOC: 0 Helbing, 1 Guo, 2 Moussaid.
// The body-compression and sliding-friction terms (k1 max(0,rd), k2 max(0,rd) dv) are identically zero unless the two bodies
// overlap, which is rare.  pair_eval<CONTACT = false> leaves them (and the relative velocity they need) out and is
// completely branch-free, so several independent evaluations interleave in the pipes; it returns rd = r_ij - d_ij, and callers
// re-evaluate with CONTACT = true only when some lane of the warp has rd > 0.  For a lane without contact both forms give the
// same bits (the extra terms enter as fma(0, ., f)).
// Helbing / Guo are evaluated on the un-normalised separation: f = (cn / d) (dx, dy) + (ct / d) (-dy, dx), so the unit vector is
// never formed, and d itself only appears inside rd = rs1 + rs2 - d2 / d (one DFMA).
template <typename T, int SOC, bool CONTACT>
__device__ __forceinline__ T pair_eval(const Params<T> &P, const double *tbl, T x1, T y1, T vx1, T vy1, T rs1, T x2, T y2, T vx2, T vy2, T rs2,
                                       T &fx, T &fy) {
    using R = Real<T>;
    const T dx = x1 - x2, dy = y1 - y2;
    const T d2 = fma_<T>(dy, dy, fma_<T>(dx, dx, tiny_<T>()));  // self pair: (dx, dy) = (0, 0) -> zero force, no branch (see tiny_)
    const T inv = R::rsqrt_(d2);
    const T rd = fma_<T>(-d2, inv, rs1 + rs2);
    if (SOC < 2) {
        const T cn = R::exp2s_times_(fma_<T>(rd, P.kBi, P.lAi), inv, tbl);
        if (SOC == 1) {
            const T ct = R::exp2s_times_(fma_<T>(rd, P.kDi, P.lCi), inv, tbl);
            fx = fma_<T>(cn, dx, -(ct * dy));
            fy = fma_<T>(cn, dy, ct * dx);
        } else {
            fx = cn * dx; fy = cn * dy;
        }
        if (CONTACT) {
            const T prd = max0(rd) * inv;
            const T dv = ((vy2 - vy1) * dx - (vx2 - vx1) * dy) * inv;  // (v2 - v1) . t,  t = (-dy, dx) / d
            const T en = P.k1 * prd, et = P.k2 * prd * dv;
            fx = fma_<T>(-et, dy, fma_<T>(en, dx, fx));
            fy = fma_<T>(et, dx, fma_<T>(en, dy, fy));
        }
    } else {
        const T dist = d2 * inv;
        const T nx = dx * inv, ny = dy * inv;
        const T ivx = fma_<T>(P.lambda, vx1 - vx2, -nx);
        const T ivy = fma_<T>(P.lambda, vy1 - vy2, -ny);
        const T i2 = np_sq(ivx, ivy) + tiny_<T>();
        const T iinv = R::rsqrt_(i2);
        const T inorm = i2 * iinv;
        const T ix = ivx * iinv, iy = ivy * iinv;
        const T theta = bound_angle<T>(R::atan2_(ny, nx) - R::atan2_(iy, ix) + R::pi());
        const T k = sign_(theta);
        const T F = P.gamma * inorm;
        // Ei exp(-d/F) exp(-(n' F theta)^2) as ONE exponential per component (the reference multiplies two, forces.py:111-112)
        const T g = dist * R::rcp_(F);
        const T a = P.ns1 * F * theta, b = P.ns * F * theta;
        T ci = R::exp2s_(fma_<T>(-R::escale(), fma_<T>(a, a, g), P.lEi), tbl);
        T ch = k * R::exp2s_(fma_<T>(-R::escale(), fma_<T>(b, b, g), P.lEi), tbl);  // coefficients of i_ij and h_ij = (-iy, ix)
        if (CONTACT) {
            const T prd = max0(rd);
            const T dvh = np_dot(vx2 - vx1, vy2 - vy1, -iy, ix);
            ci = fma_<T>(P.k1, prd, ci);
            ch = fma_<T>(P.k2 * prd, dvh, ch);
        }
        fx = -fma_<T>(ci, ix, ch * -iy);
        fy = -fma_<T>(ci, iy, ch * ix);
    }
    return rd;
}

// `vote_mask`: lanes of the warp executing this call together.
template <typename T, int SOC>
__device__ __forceinline__ void pair_force(const Params<T> &P, const double *tbl, unsigned vote_mask, T x1, T y1, T vx1, T vy1, T rs1, T x2,
                                           T y2, T vx2, T vy2, T rs2, T &fx, T &fy) {
    const T rd = pair_eval<T, SOC, false>(P, tbl, x1, y1, vx1, vy1, rs1, x2, y2, vx2, vy2, rs2, fx, fy);
    if (__any_sync(vote_mask, Real<T>::positive_(rd))) pair_eval<T, SOC, true>(P, tbl, x1, y1, vx1, vy1, rs1, x2, y2, vx2, vy2, rs2, fx, fy);
}

#ifndef SNP_SEG_UNROLL
#define SNP_SEG_UNROLL 4  // measured (4096 x 25, 14 segments): fp32 1 / 2 / 4 -> 0.1257 / 0.1273 / 0.1245 ms, fp64 0.2231 / 0.2214 / 0.2218
#endif
constexpr int kSegUnroll = SNP_SEG_UNROLL;  // wall-segment search loop (closest_point_impl)

// One wall-segment slot staged in shared memory: a, -e = a - b, e / |e|^2.  ax is NaN for padding slots.  The parameter of the
// foot point is t = (p - a) . (e / |e|^2) and p - h = (p - a) + t (-e): no sign flip and no separate scaling inside the search
// loop.  Six words, 16-byte aligned: a segment is fetched with three LDS.128 (fp64) / LDS.64 + LDS.128 (fp32).
template <typename T> struct alignas(16) Seg { T ax, ay, nex, ney, tx, ty; };

template <typename T> __device__ __forceinline__ Seg<T> make_seg(T ax, T ay, T bx, T by) {
    Seg<T> s;
    s.ax = ax; s.ay = ay; s.nex = ax - bx; s.ney = ay - by;
    const T len = np_norm(s.nex, s.ney);
    const T il2 = -Real<T>::rcp_(len * len);
    s.tx = s.nex * il2; s.ty = s.ney * il2;
    return s;
}

// Closest point of one polygon (obstacle.py:53-66): serial keeps the LAST segment among ties ('<=', init 10000),
// Numba keeps the first (np.argmin, fp:252).  Distances are compared squared (sqrt is monotone; one sqrt per polygon is
// then taken by the caller instead of one per segment).  Padding slots (ax = NaN) sit at the end of each polygon's slots
// (motion_model_manager.py:270-275) and `cnt` excludes them.
template <typename T, bool FIRST_WINS>
__device__ __forceinline__ void closest_point_impl(const Seg<T> *segs, int cnt, T px, T py, T &dxb, T &dyb, T &best) {
    best = FIRST_WINS ? Real<T>::inf() : T(1.0e8);
    dxb = px; dyb = py;  // closest point (0,0) when no segment qualifies (obstacle.py:55)
#pragma unroll kSegUnroll
    for (int s = 0; s < cnt; ++s) {
        const Seg<T> g = segs[s];
        const T qx = px - g.ax, qy = py - g.ay;
        const T t = Real<T>::clamp01_(np_dot(qx, qy, g.tx, g.ty));
        const T ux = fma_<T>(t, g.nex, qx), uy = fma_<T>(t, g.ney, qy);  // p - h,  h = a + t e
        const T d = np_sq(ux, uy);
        const bool take = FIRST_WINS ? (d < best) : (d <= best);
        best = take ? d : best; dxb = take ? ux : dxb; dyb = take ? uy : dyb;
    }
}

template <typename T>
__device__ __forceinline__ void closest_point(const Seg<T> *segs, int cnt, T px, T py, bool first_wins, T &dxb, T &dyb, T &best) {
    if (first_wins) closest_point_impl<T, true>(segs, cnt, px, py, dxb, dyb, best);   // Numba: np.argmin (fp:252)
    else closest_point_impl<T, false>(segs, cnt, px, py, dxb, dyb, best);             // serial: '<=' keeps the last (obstacle.py:63)
}

// Wall force of G consecutive polygons on one agent, accumulated into (fx, fy) in polygon order.  The G closest-point searches run
// first; the G force evaluations that follow (rsqrt -> exponent -> table -> polynomial: ~30 dependent FP64 instructions each) are
// independent and branch-free, so their chains interleave -- in a kernel that is bound by dependency latency this is worth more than
// the instructions it saves.  ONE contact vote covers the group.
template <typename T, int OBS, int G>
__device__ __forceinline__ void wall_forces(const Params<T> &P, const double *tbl, unsigned vote_mask, const Seg<T> *segs, const int *seg_cnt,
                                            int S, bool numba, T px, T py, T vx, T vy, T rs, T &fx, T &fy) {
    using R = Real<T>;
    T dx[G], dy[G], inv[G], rd[G], cn[G];
    bool touch = false;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        T d2;
        closest_point<T>(segs + k * S, seg_cnt[k], px, py, numba, dx[k], dy[k], d2);
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const T d2 = np_sq(dx[k], dy[k]) + tiny_<T>();
        inv[k] = R::rsqrt_(d2);
        rd[k] = fma_<T>(-d2, inv[k], rs);
        touch |= R::positive_(rd[k]);
        cn[k] = R::exp2s_times_(fma_<T>(rd[k], P.kBw, P.lAw), inv[k], tbl);
    }
    const bool contact = __any_sync(vote_mask, touch);  // compression / friction terms only when some lane touches a wall of the group
    if (OBS == 0 && !contact) {
#pragma unroll
        for (int k = 0; k < G; ++k) { fx = fma_<T>(cn[k], dx[k], fx); fy = fma_<T>(cn[k], dy[k], fy); }
    } else {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const T prd = contact ? max0(rd[k]) * inv[k] : T(0);
            // delta_v = -(v . t) = (vx dy - vy dx) / d,  t = (-dy, dx) / d
            const T cr = vx * dy[k] - vy * dx[k];
            T ct = -(P.k2 * prd) * cr * inv[k];
            if (OBS == 1) ct = fma_<T>(-R::exp2s_times_(fma_<T>(rd[k], P.kDw, P.lCw), inv[k] * inv[k], tbl), cr, ct);
            const T en = fma_<T>(P.k1, prd, cn[k]);
            fx += fma_<T>(en, dx[k], -(ct * dy[k]));
            fy += fma_<T>(en, dy[k], ct * dx[k]);
        }
    }
}

// Wall force of all W polygons on one agent (forces.py:27-53).  OBS: 0 Helbing (mean over walls), 1 Guo (sum; mean in Numba).
// Same un-normalised form as pair_eval: f = (cn / d) (dx, dy) + (ct / d) (-dy, dx).
#ifndef SNP_WALL_GROUP
#define SNP_WALL_GROUP 3
#endif
template <typename T, int OBS>
__device__ __forceinline__ void obstacle_force(const Params<T> &P, const double *tbl, unsigned vote_mask, const Seg<T> *segs, const int *seg_cnt,
                                               int W, int S, bool numba, T px, T py, T vx, T vy, T rs, T &fx, T &fy) {
    using R = Real<T>;
    constexpr int G = SNP_WALL_GROUP;
    fx = T(0); fy = T(0);
    int w = 0;
    if constexpr (G > 1) {
        for (; w + G <= W; w += G) wall_forces<T, OBS, G>(P, tbl, vote_mask, segs + w * S, seg_cnt + w, S, numba, px, py, vx, vy, rs, fx, fy);
    }
    for (; w < W; ++w) wall_forces<T, OBS, 1>(P, tbl, vote_mask, segs + w * S, seg_cnt + w, S, numba, px, py, vx, vy, rs, fx, fy);
    if (W > 0 && (OBS == 0 || numba)) { const T iw = R::rcp_(T(W)); fx *= iw; fy *= iw; }
}

// Per-lane agent state kept in registers across the fused sub-steps.
template <typename T> struct Agent {
    T px, py, vx, vy, th, bvx, bvy, om, dfx, dfy;  // dynamic
    T r, m, vd, rs;                                // static (rs = r + safety)
    T inv_m, mr;                                   // static derived: 1/m, m/relax_t
    T gx, gy;                                      // current goal
    T cs, sn;                                      // cos/sin(th) (headed models)
};

template <typename T> __device__ __forceinline__ void agent_static(const Params<T> &P, Agent<T> &a) {
    a.inv_m = Real<T>::rcp_(a.m);
    a.mr = a.m * P.inv_relax;
}

template <typename T> __device__ __forceinline__ void clip_speed(T &vx, T &vy, T lim) {  // mmm:52-55
    const T n2 = np_sq(vx, vy);
    if (n2 > lim * lim) {  // |v| > vd  (vd >= 0; the squares compare identically up to one rounding of vd^2)
        const T s = lim * Real<T>::rsqrt_(n2 + tiny_<T>());
        vx *= s; vy *= s;
    }
}

// Goal vector of an agent: (dx, dy) to the current goal, 1 / distance and the distance.  Shared by the goal switch
// (mmm:66-70) and the desired force (forces.py:9-16), which the reference evaluates at the same position.
template <typename T> struct GoalVec { T dx, dy, inv, dist; };
template <typename T> __device__ __forceinline__ GoalVec<T> goal_vec(const Agent<T> &a) {
    GoalVec<T> g;
    g.dx = a.gx - a.px; g.dy = a.gy - a.py;
    const T d2 = np_sq(g.dx, g.dy) + tiny_<T>();
    g.inv = Real<T>::rsqrt_(d2);
    g.dist = d2 * g.inv;
    return g;
}

// Desired force (forces.py:9-16): refreshed only outside the goal radius; inside, the serial path keeps the previous
// value (stale) while the Numba path returns zero (fp:34-40).
template <typename T> __device__ __forceinline__ void desired_force(const Params<T> &P, Agent<T> &a, const GoalVec<T> &g, bool numba) {
    if (g.dist > a.r) {
        a.dfx = a.mr * fma_<T>(g.dx * g.inv, a.vd, -a.vx);
        a.dfy = a.mr * fma_<T>(g.dy * g.inv, a.vd, -a.vy);
    } else if (numba) {
        a.dfx = T(0); a.dfy = T(0);
    }
}
template <typename T> __device__ __forceinline__ void desired_force(const Params<T> &P, Agent<T> &a, bool numba) {
    desired_force<T>(P, a, goal_vec<T>(a), numba);
}

// utils.py:7-13 for the sum / difference of two angles that are themselves bounded: |a| < 2 pi except when it is exactly
// +-2 pi, so the modulo branches of bound_angle collapse into one rarely taken test.
template <typename T> __device__ __forceinline__ T bound_angle_near(T a) {
    const T pi = Real<T>::pi(), two_pi = T(2) * pi;
    if (!(a < two_pi && a > -two_pi)) return bound_angle<T>(a);
    const T shift = a > pi ? -two_pi : (a < -pi ? two_pi : T(0));
    return a + shift;
}

// Torque (forces.py:279-290), global force (mmm:428-435) and explicit Euler (mmm:72-85) for one agent, given the wall
// force (fox, foy) and social force (fsx, fsy).  HEADED: 0 SFM, 1 HSFM torque from the desired force, 2 from the total.
// The new heading is bounded to [-pi, pi] before its sine / cosine are taken, so sincos needs no slow path.
template <typename T, int HEADED>
__device__ __forceinline__ void integrate(const Params<T> &P, Agent<T> &a, T fox, T foy, T fsx, T fsy, T dt) {
    using R = Real<T>;
    const T dtm = a.inv_m * dt;
    if (HEADED == 0) {
        const T gx = a.dfx + fox + fsx, gy = a.dfy + foy + fsy;
        a.px = fma_<T>(a.vx, dt, a.px); a.py = fma_<T>(a.vy, dt, a.py);
        a.vx = fma_<T>(gx, dtm, a.vx); a.vy = fma_<T>(gy, dtm, a.vy);
        clip_speed(a.vx, a.vy, a.vd);
    } else {
        const T ox = fox + fsx, oy = foy + fsy;
        const T sx = a.dfx + ox, sy = a.dfy + oy;
        const T tfx = HEADED == 1 ? a.dfx : sx, tfy = HEADED == 1 ? a.dfy : sy;
        // |f| = f2 * y and sqrt(|f|) = 1 / sqrt(y),  y = 1 / sqrt(f2)
        const T f2 = np_sq(tfx, tfy) + tiny_<T>();
        const T y = R::rsqrt_(f2);
        const T fn = f2 * y, sfn = R::rsqrt_(y);
        const T dth = bound_angle_near<T>(a.th - R::atan2_(tfy, tfx));
        const T aw = -(P.k_lambda * fn) * dth - (P.c_omega * sfn) * a.om;  // torque / inertia
        const T g0 = np_dot(sx, sy, a.cs, a.sn);
        const T g1 = P.ko * np_dot(ox, oy, -a.sn, a.cs) - P.kd * a.bvy;
        a.px = fma_<T>(a.vx, dt, a.px); a.py = fma_<T>(a.vy, dt, a.py);
        a.th = bound_angle_near<T>(fma_<T>(a.om, dt, a.th));
        a.bvx = fma_<T>(g0, dtm, a.bvx); a.bvy = fma_<T>(g1, dtm, a.bvy);
        a.om = fma_<T>(aw, dt, a.om);
        clip_speed(a.bvx, a.bvy, a.vd);
        R::sincos_bounded_(a.th, &a.sn, &a.cs);
        a.vx = np_mv(a.cs, -a.sn, a.bvx, a.bvy);
        a.vy = np_mv(a.sn, a.cs, a.bvx, a.bvy);
    }
}

}  // namespace snp
