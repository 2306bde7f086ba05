// snp_physics.cuh -- the SFM / HSFM force laws and the explicit-Euler update, written once and shared by the
// small-crowd fused kernel, the CTA-per-env kernel and the large-crowd tiled kernel.
//
// Reference (paths under social_gym/): src/forces.py:9-16 desired, :27-53 obstacle (Helbing / Guo), :63-128 pair laws
// (Helbing / Guo / Moussaid), :279-290 torque (Farina / "new"); src/motion_model_manager.py:424-435 global force,
// :72-85 Euler, :52-55 speed clip; src/obstacle.py:53-66 closest point on a polygon's segments.
#pragma once
#include "snp_math.cuh"

namespace snp {

// Parameter row of agent.py:269 with the reciprocals the kernels actually multiply by.
template <typename T> struct Params {
    T inv_relax, Ai, Aw, inv_Bi, inv_Bw, Ci, Cw, inv_Di, inv_Dw, Ei, k1, k2, lambda, gamma, ns, ns1, ko, kd, inv_alpha, alpha1, k_lambda;
};

template <typename T> __host__ __device__ inline Params<T> make_params(const double *p) {
    Params<T> q;
    q.inv_relax = T(1.0 / p[0]);
    q.Ai = T(p[1]); q.Aw = T(p[2]);
    q.inv_Bi = T(p[3] != 0.0 ? 1.0 / p[3] : 0.0); q.inv_Bw = T(p[4] != 0.0 ? 1.0 / p[4] : 0.0);
    q.Ci = T(p[5]); q.Cw = T(p[6]);
    q.inv_Di = T(p[7] != 0.0 ? 1.0 / p[7] : 0.0); q.inv_Dw = T(p[8] != 0.0 ? 1.0 / p[8] : 0.0);
    q.Ei = T(p[9]); q.k1 = T(p[10]); q.k2 = T(p[11]); q.lambda = T(p[12]); q.gamma = T(p[13]); q.ns = T(p[14]); q.ns1 = T(p[15]);
    q.ko = T(p[16]); q.kd = T(p[17]);
    q.inv_alpha = T(p[18] != 0.0 ? 1.0 / p[18] : 0.0); q.alpha1 = T(1.0 + p[18]); q.k_lambda = T(p[19]);
    return q;
}

// Force exerted on agent 1 by agent 2 (forces.py:63-128).  rs = radius + safety_space.
// SOC: 0 Helbing, 1 Guo, 2 Moussaid.
// The body-compression and sliding-friction terms (k1 max(0,rd), k2 max(0,rd) dv) are identically zero unless the two bodies
// overlap, which is rare.  pair_eval<CONTACT = false> leaves them (and the relative-velocity projection they need) out and is
// completely branch-free, so several independent evaluations interleave in the pipes; it returns rd = r_ij - d_ij, and callers
// re-evaluate with CONTACT = true only when some lane of the warp has rd > 0.  For a lane without contact both forms give the
// same bits (the extra terms are exact zeros added to / multiplied into the rest).
template <typename T, int SOC, bool CONTACT>
__device__ __forceinline__ T pair_eval(const Params<T> &P, const double *tbl, T x1, T y1, T vx1, T vy1, T rs1, T x2, T y2, T vx2, T vy2, T rs2,
                                       T &fx, T &fy) {
    using R = Real<T>;
    const T dx = x1 - x2, dy = y1 - y2;
    const T d2 = fma_<T>(dy, dy, fma_<T>(dx, dx, tiny_<T>()));  // self pair: n = (0,0) -> zero force, no branch (see tiny_)
    const T inv = R::rsqrt_(d2);
    const T dist = d2 * inv;
    const T nx = dx * inv, ny = dy * inv;
    const T rd = (rs1 + rs2) - dist;
    if (SOC < 2) {
        T cn = P.Ai * R::exp_(rd * P.inv_Bi, tbl);
        T ct = T(0);
        if (CONTACT) {
            const T prd = max0(rd);
            const T dv = np_dot(vx2 - vx1, vy2 - vy1, -ny, nx);  // t = (-ny, nx); dv = (v2 - v1) . t
            cn = fma_<T>(P.k1, prd, cn);
            ct = P.k2 * prd * dv;
        }
        if (SOC == 1) ct = fma_<T>(P.Ci, R::exp_(rd * P.inv_Di, tbl), ct);
        if (SOC == 1 || CONTACT) {
            fx = fma_<T>(cn, nx, ct * -ny);
            fy = fma_<T>(cn, ny, ct * nx);
        } else {
            fx = cn * nx; fy = cn * ny;
        }
    } else {
        const T ivx = fma_<T>(P.lambda, vx1 - vx2, -nx);
        const T ivy = fma_<T>(P.lambda, vy1 - vy2, -ny);
        const T i2 = np_sq(ivx, ivy) + tiny_<T>();
        const T iinv = R::rsqrt_(i2);
        const T inorm = i2 * iinv;
        const T ix = ivx * iinv, iy = ivy * iinv;
        const T theta = bound_angle<T>(R::atan2_(ny, nx) - R::atan2_(iy, ix) + R::pi());
        const T k = sign_(theta);
        const T F = P.gamma * inorm;
        const T e0 = P.Ei * R::exp_(-dist * R::rcp_(F), tbl);
        const T a = P.ns1 * F * theta, b = P.ns * F * theta;
        const T ea = R::exp_(-(a * a), tbl), eb = k * R::exp_(-(b * b), tbl);
        T ci = e0 * ea, ch = e0 * eb;  // coefficients of i_ij and h_ij = (-iy, ix)
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
    if (__any_sync(vote_mask, rd > T(0))) pair_eval<T, SOC, true>(P, tbl, x1, y1, vx1, vy1, rs1, x2, y2, vx2, vy2, rs2, fx, fy);
}

// One wall-segment slot staged in shared memory: a, -e = a - b, -1/|e|^2.  ax is NaN for padding slots.  (Negated so that
// t = ((p-a).(-e)) * (-1/|e|^2) and p - h = (p-a) + t (-e) need no sign flip inside the search loop: same bits, one DADD less.)
// Padded to six words and 16-byte aligned so that a segment is fetched with three LDS.128 (fp64) / LDS.64 + LDS.128 (fp32)
// instead of five scalar loads.
template <typename T> struct alignas(16) Seg { T ax, ay, nex, ney, ninv_len2, pad; };

template <typename T> __device__ __forceinline__ Seg<T> make_seg(T ax, T ay, T bx, T by) {
    Seg<T> s;
    s.ax = ax; s.ay = ay; s.nex = ax - bx; s.ney = ay - by;
    const T len = np_norm(s.nex, s.ney);
    s.ninv_len2 = -Real<T>::rcp_(len * len);
    s.pad = T(0);
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
    for (int s = 0; s < cnt; ++s) {
        const Seg<T> g = segs[s];
        const T qx = px - g.ax, qy = py - g.ay;
        T t = np_dot(qx, qy, g.nex, g.ney) * g.ninv_len2;
        t = max0(t);
        t = t < T(1) ? t : T(1);
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

// Wall force of all W polygons on one agent (forces.py:27-53).  OBS: 0 Helbing (mean over walls), 1 Guo (sum; mean in Numba).
template <typename T, int OBS>
__device__ __forceinline__ void obstacle_force(const Params<T> &P, const double *tbl, unsigned vote_mask, const Seg<T> *segs, const int *seg_cnt,
                                               int W, int S, bool numba, T px, T py, T vx, T vy, T rs, T &fx, T &fy) {
    using R = Real<T>;
    fx = T(0); fy = T(0);
    for (int w = 0; w < W; ++w) {
        T dx, dy, d2;
        closest_point<T>(segs + w * S, seg_cnt[w], px, py, numba, dx, dy, d2);
        d2 = np_sq(dx, dy) + tiny_<T>();
        const T inv = R::rsqrt_(d2);
        const T dist = d2 * inv;
        const T nx = dx * inv, ny = dy * inv;
        const T rd = rs - dist;
        const bool contact = __any_sync(vote_mask, rd > T(0));  // compression / friction terms only when some lane touches the wall
        T cn = P.Aw * R::exp_(rd * P.inv_Bw, tbl);
        if (OBS == 0 && !contact) {
            fx = fma_<T>(cn, nx, fx); fy = fma_<T>(cn, ny, fy);
        } else {
            const T prd = contact ? max0(rd) : T(0);
            const T dv = -np_dot(vx, vy, -ny, nx);
            cn = fma_<T>(P.k1, prd, cn);
            T ct;
            if (OBS == 0) ct = -(P.k2 * prd * dv);
            else ct = (-P.Cw * R::exp_(rd * P.inv_Dw, tbl) - P.k2 * prd) * dv;
            fx += fma_<T>(cn, nx, ct * -ny);
            fy += fma_<T>(cn, ny, ct * nx);
        }
    }
    if (W > 0 && (OBS == 0 || numba)) { const T iw = R::rcp_(T(W)); fx *= iw; fy *= iw; }
}

// Per-lane agent state kept in registers across the fused sub-steps.
template <typename T> struct Agent {
    T px, py, vx, vy, th, bvx, bvy, om, dfx, dfy;  // dynamic
    T r, m, vd, rs;                                // static (rs = r + safety)
    T inv_m, mr, inertia, inv_inertia;             // static derived: 1/m, m/relax_t, 0.5 m r^2 (agent.py:30) and its inverse
    T gx, gy;                                      // current goal
    T cs, sn;                                      // cos/sin(th) (headed models)
};

template <typename T> __device__ __forceinline__ void agent_static(const Params<T> &P, Agent<T> &a) {
    a.inv_m = Real<T>::rcp_(a.m);
    a.mr = a.m * P.inv_relax;
    a.inertia = T(0.5) * a.m * a.r * a.r;
    a.inv_inertia = Real<T>::rcp_(a.inertia);
}

template <typename T> __device__ __forceinline__ void clip_speed(T &vx, T &vy, T lim) {  // mmm:52-55
    const T n2 = np_sq(vx, vy);
    if (n2 > lim * lim) {  // |v| > vd  (vd >= 0; the squares compare identically up to one rounding of vd^2)
        const T s = lim * Real<T>::rsqrt_(n2 + tiny_<T>());
        vx *= s; vy *= s;
    }
}

// Desired force (forces.py:9-16): refreshed only outside the goal radius; inside, the serial path keeps the previous
// value (stale) while the Numba path returns zero (fp:34-40).
template <typename T> __device__ __forceinline__ void desired_force(const Params<T> &P, Agent<T> &a, bool numba) {
    const T dx = a.gx - a.px, dy = a.gy - a.py;
    const T d2 = np_sq(dx, dy) + tiny_<T>();
    const T inv = Real<T>::rsqrt_(d2);
    const T dist = d2 * inv;
    if (dist > a.r) {
        a.dfx = a.mr * fma_<T>(dx * inv, a.vd, -a.vx);
        a.dfy = a.mr * fma_<T>(dy * inv, a.vd, -a.vy);
    } else if (numba) {
        a.dfx = T(0); a.dfy = T(0);
    }
}

// Torque (forces.py:279-290), global force (mmm:428-435) and explicit Euler (mmm:72-85) for one agent, given the wall
// force (fox, foy) and social force (fsx, fsy).  HEADED: 0 SFM, 1 HSFM torque from the desired force, 2 from the total.
template <typename T, int HEADED>
__device__ __forceinline__ void integrate(const Params<T> &P, Agent<T> &a, T fox, T foy, T fsx, T fsy, T dt) {
    using R = Real<T>;
    const T inv_m = a.inv_m;
    if (HEADED == 0) {
        const T gx = a.dfx + fox + fsx, gy = a.dfy + foy + fsy;
        a.px = fma_<T>(a.vx, dt, a.px); a.py = fma_<T>(a.vy, dt, a.py);
        a.vx = fma_<T>(gx * inv_m, dt, a.vx); a.vy = fma_<T>(gy * inv_m, dt, a.vy);
        clip_speed(a.vx, a.vy, a.vd);
    } else {
        const T sx = a.dfx + fox + fsx, sy = a.dfy + foy + fsy;
        const T tfx = HEADED == 1 ? a.dfx : sx, tfy = HEADED == 1 ? a.dfy : sy;
        const T inertia = a.inertia;
        const T fn = np_norm(tfx, tfy);
        const T k_theta = inertia * P.k_lambda * fn;
        const T k_omega = inertia * P.alpha1 * R::sqrt_(P.k_lambda * fn * P.inv_alpha);
        const T tq = -k_theta * bound_angle<T>(a.th - R::atan2_(tfy, tfx)) - k_omega * a.om;
        const T g0 = np_dot(sx, sy, a.cs, a.sn);
        const T g1 = P.ko * np_dot(fox + fsx, foy + fsy, -a.sn, a.cs) - P.kd * a.bvy;
        a.px = fma_<T>(a.vx, dt, a.px); a.py = fma_<T>(a.vy, dt, a.py);
        a.th = bound_angle<T>(fma_<T>(a.om, dt, a.th));
        a.bvx = fma_<T>(g0 * inv_m, dt, a.bvx); a.bvy = fma_<T>(g1 * inv_m, dt, a.bvy);
        a.om = fma_<T>(tq * a.inv_inertia, dt, a.om);
        clip_speed(a.bvx, a.bvy, a.vd);
        R::sincos_(a.th, &a.sn, &a.cs);
        a.vx = np_mv(a.cs, -a.sn, a.bvx, a.bvy);
        a.vy = np_mv(a.sn, a.cs, a.bvx, a.bvy);
    }
}

}  // namespace snp
