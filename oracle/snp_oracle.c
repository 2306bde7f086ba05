/*
 * snp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, fp64, operation-for-operation restatement of the reference's serial
 * Python/NumPy human motion update and the checks either side of it.  It exists so that
 *   - tests/ can check the CUDA path against the reference's algorithm at sizes the Python
 *     reference cannot reach (4096 envs, 65k agents),
 *   - bench.py can time a CPU baseline (`cpu_baseline`, `--impl reference`) on the GPU box,
 *     where /root/reference does not exist.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (social_navigation_pyenvs_b200/) never does.
 *
 * PARITY PIN: this file is validated against golden vectors produced by running the live
 * reference (tests/golden/make_golden.py writes the .npz files under tests/golden/; tests/test_oracle_golden.py).
 * The reference ships no tests of its own, so those recorded runs of the
 * reference itself are the pin.
 *
 * Reference citations use paths relative to /root/reference/social_gym/:
 *   mmm = src/motion_model_manager.py, forces = src/forces.py, fp = src/forces_parallel.py,
 *   sim = social_nav_sim.py, gym = social_nav_gym.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC snp_oracle.c -lm   (see oracle/build.py)
 * -ffp-contract=off keeps gcc from fusing a*b+c, which CPython/NumPy scalar code never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NS 13 /* state row: px,py,theta,vx,vy,bvx,bvy,omega,r,m,gx,gy,vd  (src/agent.py:256-258) */
#define NP 20 /* params row (src/agent.py:269) */

enum { P_RELAX = 0, P_AI, P_AW, P_BI, P_BW, P_CI, P_CW, P_DI, P_DW, P_EI, P_K1, P_K2, P_LAMBDA, P_GAMMA, P_NS, P_NS1, P_KO, P_KD, P_ALPHA, P_KLAMBDA };

typedef struct {
    int type;           /* 0..8, index into SFMS (mmm:15-17) */
    int n;              /* humans per env */
    int g;              /* goal slots per human (NaN padded) */
    int n_walls;        /* wall polygons */
    int n_segs;         /* segment slots per wall (NaN padded) */
    int consider_robot; /* robot row at index n exerts force (mmm:35,259) */
    int symmetric;      /* all_equal_humans -> compute_all_social_forces (mmm:455-457) */
    int numba_compat;   /* 0: serial semantics (the oracle of record); 1: forces_parallel.py semantics */
    int walls_per_env;  /* 0: one wall set shared by all envs; 1: walls[E][W][S][2][2] */
    int respawn;        /* parallel-traffic respawn after every update (mmm:407-422) */
    double respawn_bounds[2]; /* (traffic_length / 2, traffic_height / 2)  (sim:360) */
} orc_cfg;

/* src/utils.py:7-13 (Python float %: result takes the sign of the divisor; here dividend and
 * divisor always share a sign, so C fmod gives the same value). */
static double bound_angle(double a) {
    const double two_pi = 2.0 * M_PI;
    if (a >= two_pi) a = fmod(a, two_pi);
    if (a <= -two_pi) a = fmod(a, -two_pi);
    if (a > M_PI) a -= two_pi;
    if (a < -M_PI) a += two_pi;
    return a;
}

/* The reference mixes two norm/dot flavours and they differ in the last bit:
 *  - utils.py:42-48 two_dim_norm / two_dim_dot_product (also everything inside Numba): plain x*x + y*y;
 *  - np.linalg.norm / np.dot / np.matmul on length-2 float64 arrays: NumPy hands these to OpenBLAS, whose
 *    x86-64 FMA kernels evaluate  fma(a1, b1, a0*b0)  (ddot) and  fma(R_i0, b0, R_i1*b1)  (gemv) -- measured
 *    against NumPy 2.3.5 / OpenBLAS 0.3.30 on 20000 random inputs, 100% bit-identical (tests/test_oracle_golden.py).
 * fm = 1 selects the NumPy/OpenBLAS flavour (serial path), fm = 0 the plain one (Numba path). */
static inline double norm2(double x, double y) { return sqrt(x * x + y * y); }
static inline double np_norm(int fm, double x, double y) { return fm ? sqrt(fma(y, y, x * x)) : sqrt(x * x + y * y); }
static inline double np_dot(int fm, double a0, double a1, double b0, double b1) { return fm ? fma(a1, b1, a0 * b0) : a0 * b0 + a1 * b1; }
static inline double np_mv(int fm, double r0, double r1, double b0, double b1) { return fm ? fma(r0, b0, r1 * b1) : r0 * b0 + r1 * b1; }
static inline double pos(double x) { return x > 0.0 ? x : 0.0; } /* max(0, x) */
static inline double sgn(double x) { return (x > 0.0) - (x < 0.0); }

/* forces.py:63-128 compute_pairwise_social_force, soc in {0 Helbing, 1 Guo, 2 Moussaid}.
 * a1 = repulsed agent (its params are used), a2 = the agent exerting the force. */
static void pair_force(int soc, int fm, const double *a1, double s1, const double *a2, double s2, const double *p, double *f) {
    double r_ij = a1[8] + s1 + a2[8] + s2;                   /* forces.py:80 */
    double dx = a1[0] - a2[0], dy = a1[1] - a2[1];           /* :81 */
    double dist = np_norm(fm, dx, dy);                       /* :82 */
    double nx = dx / dist, ny = dy / dist;                   /* :83 */
    double rd = r_ij - dist;                                 /* :84 */
    if (soc == 0) {                                          /* :85-88 */
        double tx = -ny, ty = nx;
        double dv = np_dot(fm, a2[3] - a1[3], a2[4] - a1[4], tx, ty);
        double cn = p[P_AI] * exp(rd / p[P_BI]) + p[P_K1] * pos(rd);
        double ct = p[P_K2] * pos(rd) * dv;
        f[0] = cn * nx + ct * tx;
        f[1] = cn * ny + ct * ty;
    } else if (soc == 1) {                                   /* :89-93 */
        double tx = -ny, ty = nx;
        double dv = np_dot(fm, a2[3] - a1[3], a2[4] - a1[4], tx, ty);
        double cn = p[P_AI] * exp(rd / p[P_BI]) + p[P_K1] * pos(rd);
        double ct = p[P_CI] * exp(rd / p[P_DI]) + p[P_K2] * pos(rd) * dv;
        f[0] = cn * nx + ct * tx;
        f[1] = cn * ny + ct * ty;
    } else {                                                 /* :96-110 */
        double ivx = p[P_LAMBDA] * (a1[3] - a2[3]) - nx;
        double ivy = p[P_LAMBDA] * (a1[4] - a2[4]) - ny;
        double inorm = np_norm(fm, ivx, ivy);
        double ix = ivx / inorm, iy = ivy / inorm;
        double theta = bound_angle(atan2(ny, nx) - atan2(iy, ix) + M_PI);
        double k = sgn(theta);
        double hx = -iy, hy = ix;
        double F = p[P_GAMMA] * inorm;
        double dvh = np_dot(fm, a2[3] - a1[3], a2[4] - a1[4], hx, hy);
        double e0 = p[P_EI] * exp(-dist / F);
        double a = p[P_NS1] * F * theta, b = p[P_NS] * F * theta;
        double ea = exp(-(a * a)), eb = k * exp(-(b * b));
        double c1 = p[P_K1] * pos(rd), c2 = p[P_K2] * pos(rd) * dvh;
        f[0] = -(e0 * (ea * ix + eb * hx) + c1 * ix + c2 * hx);
        f[1] = -(e0 * (ea * iy + eb * hy) + c1 * iy + c2 * hy);
    }
}

/* src/obstacle.py:53-66 get_closest_point (serial: ties -> last segment, init 10000 / (0,0));
 * fp:236-252 (Numba: NaN slots get INT64_MAX, np.argmin -> first). */
static void closest_point(const double *wall, int n_segs, double px, double py, int numba, double *cx, double *cy) {
    const int fm = !numba;
    double best = numba ? INFINITY : 10000.0;
    double bx = 0.0, by = 0.0;
    int have = 0;
    for (int s = 0; s < n_segs; ++s) {
        const double *sg = wall + 4 * s;
        double d, hx, hy;
        if (isnan(sg[0])) {
            if (!numba) continue;
            d = (double)INT64_MAX; hx = 0.0; hy = 0.0;       /* fp:245-246 (closest_points stays zero) */
        } else {
            double ax = sg[0], ay = sg[1], ex = sg[2] - sg[0], ey = sg[3] - sg[1];
            double len = np_norm(fm, ex, ey);
            double t = np_dot(fm, px - ax, py - ay, ex, ey) / (len * len);
            double ts = t > 0.0 ? t : 0.0;                   /* min(max(0,t),1) */
            ts = ts < 1.0 ? ts : 1.0;
            hx = ax + ts * ex; hy = ay + ts * ey;
            d = np_norm(fm, hx - px, hy - py);
        }
        if (numba ? (!have || d < best) : (d <= best)) { best = d; bx = hx; by = hy; have = 1; }
    }
    *cx = bx; *cy = by;
}

/* forces.py:27-37 (Helbing, /W) and :39-53 (Guo, no /W); fp:135-162 divides both by W. */
static void obstacle_force(int obs, int numba, const double *a, double s, const double *p, const double *cp, int W, double *f) {
    double fx = 0.0, fy = 0.0;
    const int fm = !numba;
    for (int w = 0; w < W; ++w) {
        double dx = a[0] - cp[2 * w], dy = a[1] - cp[2 * w + 1];
        double dist = np_norm(fm, dx, dy);
        double nx = dx / dist, ny = dy / dist, tx = -ny, ty = nx;
        double dv = -np_dot(fm, a[3], a[4], tx, ty);
        double rd = numba ? (a[8] - dist + s) : (a[8] + s - dist);
        double cn = p[P_AW] * exp(rd / p[P_BW]) + p[P_K1] * pos(rd);
        if (obs == 0) {
            double ct = p[P_K2] * pos(rd) * dv;
            fx += cn * nx - ct * tx;
            fy += cn * ny - ct * ty;
        } else {
            double ct = (-p[P_CW] * exp(rd / p[P_DW]) - p[P_K2] * pos(rd)) * dv;
            fx += cn * nx + ct * tx;
            fy += cn * ny + ct * ty;
        }
    }
    if (W > 0 && (obs == 0 || numba)) { fx /= W; fy /= W; }
    f[0] = fx; f[1] = fy;
}

/* forces.py:279-290 / fp:164-182 */
static double torque_force(int fm, const double *a, double inertia, double fx, double fy, const double *p) {
    double fn = np_norm(fm, fx, fy);
    double k_theta = inertia * p[P_KLAMBDA] * fn;
    double k_omega = inertia * (1 + p[P_ALPHA]) * sqrt((p[P_KLAMBDA] * fn) / p[P_ALPHA]);
    return -k_theta * bound_angle(a[2] - atan2(fy, fx)) - k_omega * a[7];
}

static void clip_norm(int fm, double *vx, double *vy, double lim) { /* mmm:52-55 */
    double n = np_norm(fm, *vx, *vy);
    if (n > lim) { *vx = (*vx / n) * lim; *vy = (*vy / n) * lim; }
}

/* One env, one update_humans call (mmm:354-373 serial Euler path; Appendix B of SURVEY.md).
 * st: [n+R][13] updated in place (rows < n); goals: [n][g][2] rotated in place;
 * desired: [n][2] carried desired force (serial path leaves it stale inside the goal radius,
 * forces.py:12-15); forces_out: optional [n][9] = desired, obstacle, social, torque, global. */
static void update_env(const orc_cfg *c, double *st, double *goals, const double *walls, const double *params,
                       const double *safety, double *desired, double dt, double *forces_out, double *scratch) {
    const int n = c->n, W = c->n_walls, ents = n + (c->consider_robot ? 1 : 0);
    const int soc = c->type % 3;
    const int obs = (c->type == 1 || c->type == 4 || c->type == 7) ? 1 : 0;
    const int headed = c->type / 3;
    const int fm = !c->numba_compat;
    double *cp = scratch;                 /* [n][W][2] */
    double *soc_f = cp + (size_t)n * (W > 0 ? W : 1) * 2; /* [n][2] */
    double *rot = soc_f + 2 * n;          /* [n][2] cos,sin */
    /* ---- compute_forces first loop (mmm:439-453) ---- */
    for (int i = 0; i < n; ++i) {
        double *a = st + NS * i;
        double *gl = goals + (size_t)i * c->g * 2;
        double dg = np_norm(fm, gl[0] - a[0], gl[1] - a[1]);
        if (c->numba_compat ? (dg <= a[8]) : (dg < a[8])) { /* mmm:67 vs fp:226 */
            int cnt = 0;
            while (cnt < c->g && !isnan(gl[2 * cnt])) ++cnt;
            double g0 = gl[0], g1 = gl[1];
            for (int k = 0; k + 1 < cnt; ++k) { gl[2 * k] = gl[2 * k + 2]; gl[2 * k + 1] = gl[2 * k + 3]; }
            if (cnt > 0) { gl[2 * (cnt - 1)] = g0; gl[2 * (cnt - 1) + 1] = g1; }
        }
        a[10] = gl[0]; a[11] = gl[1];
        for (int w = 0; w < W; ++w)
            closest_point(walls + (size_t)w * c->n_segs * 4, c->n_segs, a[0], a[1], c->numba_compat, &cp[(i * W + w) * 2], &cp[(i * W + w) * 2 + 1]);
        if (headed) { /* mmm:143-145,448 ; src/agent.py:76-77 */
            double cs = cos(a[2]), sn = sin(a[2]);
            rot[2 * i] = cs; rot[2 * i + 1] = sn;
            a[3] = np_mv(fm, cs, -sn, a[5], a[6]);
            a[4] = np_mv(fm, sn, cs, a[5], a[6]);
        }
    }
    /* ---- social forces ---- */
    memset(soc_f, 0, sizeof(double) * 2 * n);
    if (c->symmetric) { /* forces.py:130-151 (fp:86-133 sums the same terms per row) */
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < ents; ++j) {
                double f[2];
                pair_force(soc, fm, st + NS * i, safety[i], st + NS * j, safety[j], params + NP * (c->numba_compat ? 0 : i), f);
                soc_f[2 * i] += f[0]; soc_f[2 * i + 1] += f[1];
                if (j < n) { soc_f[2 * j] -= f[0]; soc_f[2 * j + 1] -= f[1]; }
            }
    } else { /* forces.py:153-218 per-agent path / fp:42-84 */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < ents; ++j) {
                if (j == i) continue;
                double f[2];
                pair_force(soc, fm, st + NS * i, safety[i], st + NS * j, safety[j], params + NP * i, f);
                soc_f[2 * i] += f[0]; soc_f[2 * i + 1] += f[1];
            }
    }
    /* ---- per-human forces (mmm:424-435) then Euler (mmm:72-85). Forces of every human are
     * computed before any state changes, as in mmm:369-373. ---- */
    double *glob = rot + 2 * n; /* [n][3] gx, gy, torque */
    for (int i = 0; i < n; ++i) {
        double *a = st + NS * i;
        const double *p = params + NP * i;
        double *df = desired + 2 * i;
        double dx = a[10] - a[0], dy = a[11] - a[1];
        double dist = np_norm(fm, dx, dy);
        if (dist > a[8]) { /* forces.py:12-14 */
            double ex = dx / dist, ey = dy / dist;
            df[0] = a[9] * (ex * a[12] - a[3]) / p[P_RELAX];
            df[1] = a[9] * (ey * a[12] - a[4]) / p[P_RELAX];
        } else if (c->numba_compat) { df[0] = 0.0; df[1] = 0.0; } /* fp:34-40 */
        double fo[2];
        obstacle_force(obs, c->numba_compat, a, safety[i], p, cp + (size_t)i * W * 2, W, fo);
        double fs0 = soc_f[2 * i], fs1 = soc_f[2 * i + 1];
        double tq = 0.0, g0, g1;
        if (!headed) { /* mmm:430 */
            g0 = df[0] + fo[0] + fs0; g1 = df[1] + fo[1] + fs1;
        } else {
            double inertia = 0.5 * a[9] * a[8] * a[8]; /* src/agent.py:30 */
            double sx = df[0] + fo[0] + fs0, sy = df[1] + fo[1] + fs1;
            tq = (headed == 1) ? torque_force(fm, a, inertia, df[0], df[1], p) : torque_force(fm, a, inertia, sx, sy, p);
            double cs = rot[2 * i], sn = rot[2 * i + 1];
            g0 = np_dot(fm, sx, sy, cs, sn);                                            /* mmm:434 */
            g1 = p[P_KO] * np_dot(fm, fo[0] + fs0, fo[1] + fs1, -sn, cs) - p[P_KD] * a[6]; /* mmm:435 */
        }
        glob[3 * i] = g0; glob[3 * i + 1] = g1; glob[3 * i + 2] = tq;
        if (forces_out) {
            double *o = forces_out + 9 * i;
            o[0] = df[0]; o[1] = df[1]; o[2] = fo[0]; o[3] = fo[1]; o[4] = fs0; o[5] = fs1; o[6] = tq; o[7] = g0; o[8] = g1;
        }
    }
    for (int i = 0; i < n; ++i) {
        double *a = st + NS * i;
        a[0] += a[3] * dt; a[1] += a[4] * dt;
        if (!headed) { /* mmm:72-76 */
            a[3] += (glob[3 * i] / a[9]) * dt; a[4] += (glob[3 * i + 1] / a[9]) * dt;
            clip_norm(fm, &a[3], &a[4], a[12]);
        } else { /* mmm:78-85 */
            double inertia = 0.5 * a[9] * a[8] * a[8];
            a[2] = bound_angle(a[2] + a[7] * dt);
            a[5] += (glob[3 * i] / a[9]) * dt; a[6] += (glob[3 * i + 1] / a[9]) * dt;
            a[7] += (glob[3 * i + 2] / inertia) * dt;
            clip_norm(fm, &a[5], &a[6], a[12]);
            double cs = cos(a[2]), sn = sin(a[2]);
            a[3] = np_mv(fm, cs, -sn, a[5], a[6]);
            a[4] = np_mv(fm, sn, cs, a[5], a[6]);
        }
    }
    /* ---- post_update: parallel-traffic respawn (mmm:407-422), sequential in index order: the max runs over the humans as
     * already moved / respawned so far.  The goal list becomes the single goal (gx, new y) (mmm:418). ---- */
    if (c->respawn) {
        for (int i = 0; i < n; ++i) {
            double *a = st + NS * i;
            double *gl = goals + (size_t)i * c->g * 2;
            if (np_norm(1, a[0] - gl[0], a[1] - gl[1]) < 3) {
                double xmax = -INFINITY, rsmax = -INFINITY;
                for (int k = 0; k < n; ++k) {
                    if (st[NS * k] > xmax) xmax = st[NS * k];
                    if (st[NS * k + 8] + safety[k] > rsmax) rsmax = st[NS * k + 8] + safety[k];
                }
                if (c->consider_robot) {
                    if (st[NS * n] > xmax) xmax = st[NS * n];
                    if (st[NS * n + 8] + safety[n] > rsmax) rsmax = st[NS * n + 8] + safety[n];
                }
                double nx = xmax + rsmax * 2;
                a[0] = nx > c->respawn_bounds[0] ? nx : c->respawn_bounds[0];
                if (a[1] >= 0) a[1] = a[1] < c->respawn_bounds[1] ? a[1] : c->respawn_bounds[1];
                else a[1] = a[1] > -c->respawn_bounds[1] ? a[1] : -c->respawn_bounds[1];
                gl[1] = a[1];
                for (int k = 1; k < c->g; ++k) { gl[2 * k] = NAN; gl[2 * k + 1] = NAN; }
                a[10] = gl[0]; a[11] = gl[1];
            }
        }
    }
}

static size_t scratch_doubles(const orc_cfg *c) { return (size_t)c->n * ((c->n_walls > 0 ? c->n_walls : 1) * 2 + 2 + 2 + 3) + 16; }

/* Batched update_humans over E independent envs, n_steps sub-steps each.
 * robot_vel (optional, [E][2]): before every sub-step the robot row moves as RobotAgent.step does for
 * a holonomic action (src/robot_agent.py:126-131): p += v*dt, linear_velocity = v. */
void orc_update_humans(const orc_cfg *c, int E, double *states, double *goals, const double *walls, const double *params,
                       const double *safety, double *desired, double dt, int n_steps, const double *robot_vel,
                       double *forces_out, int n_threads) {
    const int rows = c->n + (c->consider_robot ? 1 : 0);
    const size_t wstride = c->walls_per_env ? (size_t)c->n_walls * c->n_segs * 4 : 0;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        double *scratch = (double *)malloc(sizeof(double) * scratch_doubles(c));
#pragma omp for schedule(static)
        for (int e = 0; e < E; ++e) {
            double *st = states + (size_t)e * rows * NS;
            for (int s = 0; s < n_steps; ++s) {
                if (robot_vel && c->consider_robot) {
                    double *r = st + NS * c->n;
                    r[0] = r[0] + robot_vel[2 * e] * dt; r[1] = r[1] + robot_vel[2 * e + 1] * dt;
                    r[3] = robot_vel[2 * e]; r[4] = robot_vel[2 * e + 1];
                }
                update_env(c, st, goals + (size_t)e * c->n * c->g * 2, walls + e * wstride, params + (size_t)e * c->n * NP,
                           safety + (size_t)e * rows, desired + (size_t)e * c->n * 2, dt,
                           forces_out ? forces_out + (size_t)e * c->n * 9 : NULL, scratch);
            }
        }
        free(scratch);
    }
}

/* One update_robot call (mmm:593-653, Euler, just_velocities=False): the robot moves by its own SFM / HSFM model.
 * rb: the robot's 13-wide row, rgoals: [rg][2] goal list (rotated in place), rdes: carried desired force [2],
 * rp: the robot's 20 parameters, rtype: its model (0..8), rs: its safety space.  Humans exert force on it through the
 * per-agent path compute_social_force_*(index = len(humans), ..., consider_robot = False) (mmm:608, forces.py:153-218). */
static void update_robot_env(const orc_cfg *c, double *rb, double *rgoals, int rg, double *rdes, const double *rp, int rtype, double rs,
                             const double *st, const double *safety, const double *walls, double dt, double *scratch, int just_velocities) {
    const int n = c->n, W = c->n_walls, fm = 1;
    const int soc = rtype % 3, obs = (rtype == 1 || rtype == 4 || rtype == 7) ? 1 : 0, headed = rtype / 3;
    double *cp = scratch;
    if (np_norm(fm, rgoals[0] - rb[0], rgoals[1] - rb[1]) < rb[8]) { /* mmm:598 update_goals(robot) */
        int cnt = 0;
        while (cnt < rg && !isnan(rgoals[2 * cnt])) ++cnt;
        double g0 = rgoals[0], g1 = rgoals[1];
        for (int k = 0; k + 1 < cnt; ++k) { rgoals[2 * k] = rgoals[2 * k + 2]; rgoals[2 * k + 1] = rgoals[2 * k + 3]; }
        if (cnt > 0) { rgoals[2 * (cnt - 1)] = g0; rgoals[2 * (cnt - 1) + 1] = g1; }
    }
    rb[10] = rgoals[0]; rb[11] = rgoals[1];
    for (int w = 0; w < W; ++w) closest_point(walls + (size_t)w * c->n_segs * 4, c->n_segs, rb[0], rb[1], 0, &cp[2 * w], &cp[2 * w + 1]);
    double cs = 1.0, sn = 0.0;
    if (headed) { /* mmm:605 */
        cs = cos(rb[2]); sn = sin(rb[2]);
        rb[3] = np_mv(fm, cs, -sn, rb[5], rb[6]);
        rb[4] = np_mv(fm, sn, cs, rb[5], rb[6]);
    }
    double dx = rb[10] - rb[0], dy = rb[11] - rb[1];
    double dist = np_norm(fm, dx, dy);
    if (dist > rb[8]) {
        rdes[0] = rb[9] * ((dx / dist) * rb[12] - rb[3]) / rp[P_RELAX];
        rdes[1] = rb[9] * ((dy / dist) * rb[12] - rb[4]) / rp[P_RELAX];
    }
    double fo[2];
    obstacle_force(obs, 0, rb, rs, rp, cp, W, fo);
    double fs0 = 0.0, fs1 = 0.0;
    for (int j = 0; j < n; ++j) {
        double f[2];
        pair_force(soc, fm, rb, rs, st + NS * j, safety[j], rp, f);
        fs0 += f[0]; fs1 += f[1];
    }
    if (!just_velocities) { rb[0] += rb[3] * dt; rb[1] += rb[4] * dt; } /* mmm:73,80 */
    if (!headed) { /* mmm:609,628 */
        double g0 = rdes[0] + fo[0] + fs0, g1 = rdes[1] + fo[1] + fs1;
        rb[3] += (g0 / rb[9]) * dt; rb[4] += (g1 / rb[9]) * dt;
        clip_norm(fm, &rb[3], &rb[4], rb[12]);
    } else { /* mmm:611-613,629 */
        double inertia = 0.5 * rb[9] * rb[8] * rb[8];
        double sx = rdes[0] + fo[0] + fs0, sy = rdes[1] + fo[1] + fs1;
        double tq = (headed == 1) ? torque_force(fm, rb, inertia, rdes[0], rdes[1], rp) : torque_force(fm, rb, inertia, sx, sy, rp);
        double g0 = np_dot(fm, sx, sy, cs, sn);
        double g1 = rp[P_KO] * np_dot(fm, fo[0] + fs0, fo[1] + fs1, -sn, cs) - rp[P_KD] * rb[6];
        if (!just_velocities) rb[2] = bound_angle(rb[2] + rb[7] * dt); /* mmm:81 */
        rb[5] += (g0 / rb[9]) * dt; rb[6] += (g1 / rb[9]) * dt;
        rb[7] += (tq / inertia) * dt;
        clip_norm(fm, &rb[5], &rb[6], rb[12]);
        double c2 = cos(rb[2]), s2 = sin(rb[2]);
        rb[3] = np_mv(fm, c2, -s2, rb[5], rb[6]);
        rb[4] = np_mv(fm, s2, c2, rb[5], rb[6]);
    }
}

/* n_steps x (update_robot; update_humans): the sub-step loop of SocialNavGym.imitation_learning_step (gym:260-265).
 * robot [E][13] (updated in place; copied into row n of `states` when consider_robot), robot_goals [E][rg][2],
 * robot_desired [E][2], robot_params [20], robot_safety [E]. */
void orc_imitation_steps(const orc_cfg *c, int E, double *states, double *goals, const double *walls, const double *params,
                         const double *safety, double *desired, double dt, int n_steps, double *robot, double *robot_goals, int rg,
                         double *robot_desired, const double *robot_params, int robot_type, const double *robot_safety, int n_threads) {
    const int rows = c->n + (c->consider_robot ? 1 : 0);
    const size_t wstride = c->walls_per_env ? (size_t)c->n_walls * c->n_segs * 4 : 0;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        double *scratch = (double *)malloc(sizeof(double) * scratch_doubles(c));
#pragma omp for schedule(static)
        for (int e = 0; e < E; ++e) {
            double *st = states + (size_t)e * rows * NS;
            double *rb = robot + (size_t)e * NS;
            for (int s = 0; s < n_steps; ++s) {
                update_robot_env(c, rb, robot_goals + (size_t)e * rg * 2, rg, robot_desired + 2 * e, robot_params, robot_type, robot_safety[e],
                                 st, safety + (size_t)e * rows, walls + e * wstride, dt, scratch, 0);
                if (c->consider_robot) memcpy(st + NS * c->n, rb, sizeof(double) * NS);
                update_env(c, st, goals + (size_t)e * c->n * c->g * 2, walls + e * wstride, params + (size_t)e * c->n * NP,
                           safety + (size_t)e * rows, desired + (size_t)e * c->n * 2, dt, NULL, scratch);
            }
        }
        free(scratch);
    }
}

/* n_steps x SocialNavSim.update (sim:476-492) with the robot driven by a motion model through control_robot (sim:500-529):
 * every update the pose advances with the last velocity (update_robot_pose, mmm:655-657: yaw is NOT wrapped) and, when the update
 * index is a multiple of `every` (is_multiple(sim_t, ROBOT_SAMPLING_TIME), sim:501), the velocities are refreshed by
 * update_robot(sim_t, robot_dt, just_velocities=True) (sim:523-524); with every <= 1 (equal sampling times, sim:521)
 * update_robot(sim_t, dt) moves the robot.  The humans are then updated seeing the robot's state from BEFORE control_robot
 * (get_safe_state / set_state around it, sim:484-491; the goal list is not part of that state).  `phase` = index of the first update. */
void orc_sim_update_steps(const orc_cfg *c, int E, double *states, double *goals, const double *walls, const double *params,
                          const double *safety, double *desired, double dt, int n_steps, double *robot, double *robot_goals, int rg,
                          double *robot_desired, const double *robot_params, int robot_type, const double *robot_safety, int every,
                          double robot_dt, int phase, int n_threads) {
    const int rows = c->n + (c->consider_robot ? 1 : 0);
    const size_t wstride = c->walls_per_env ? (size_t)c->n_walls * c->n_segs * 4 : 0;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        double *scratch = (double *)malloc(sizeof(double) * scratch_doubles(c));
#pragma omp for schedule(static)
        for (int e = 0; e < E; ++e) {
            double *st = states + (size_t)e * rows * NS;
            double *rb = robot + (size_t)e * NS;
            for (int s = 0; s < n_steps; ++s) {
                double before[8], after[8];
                memcpy(before, rb, sizeof(before));
                if (every <= 1) {
                    update_robot_env(c, rb, robot_goals + (size_t)e * rg * 2, rg, robot_desired + 2 * e, robot_params, robot_type, robot_safety[e],
                                     st, safety + (size_t)e * rows, walls + e * wstride, dt, scratch, 0);
                } else {
                    rb[0] += rb[3] * dt; rb[1] += rb[4] * dt; rb[2] += rb[7] * dt;
                    if ((phase + s) % every == 0)
                        update_robot_env(c, rb, robot_goals + (size_t)e * rg * 2, rg, robot_desired + 2 * e, robot_params, robot_type,
                                         robot_safety[e], st, safety + (size_t)e * rows, walls + e * wstride, robot_dt, scratch, 1);
                }
                memcpy(after, rb, sizeof(after));
                memcpy(rb, before, sizeof(before));
                if (c->consider_robot) memcpy(st + NS * c->n, rb, sizeof(double) * NS);
                update_env(c, st, goals + (size_t)e * c->n * c->g * 2, walls + e * wstride, params + (size_t)e * c->n * NP,
                           safety + (size_t)e * rows, desired + (size_t)e * c->n * 2, dt, NULL, scratch);
                memcpy(rb, after, sizeof(after));
            }
        }
        free(scratch);
    }
}

/* utils.py:22-36 */
static double point_to_segment_dist(double x1, double y1, double x2, double y2, double x3, double y3) {
    double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return norm2(x3 - x1, y3 - y1);
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1; else if (u < 0) u = 0;
    double x = x1 + u * px, y = y1 + u * py;
    return norm2(x - x3, y - y3);
}

/* Per env: humans [n][13] rows (only px,py,vx,vy,r read), robot row [13], action [2].
 * out [12] per env:
 *  0 collision (sim:949-984 swept)   1 dmin   2 reaching_goal
 *  3 reward 4 terminated 5 truncated 6 info code (sim:986-1029; 0 Nothing 1 Timeout 2 Collision 3 ReachGoal 4 Danger)
 *  7 actual collision 8 actual dmin 9 actual goal (gym:107-118)   10 run_k_steps collision (sim:702-703)  11 unused
 * consts: time_limit, collision_penalty, success_reward, discomfort_dist, discomfort_penalty_factor, robot_time_step */
void orc_checks(int E, int n, int rows, const double *states, const double *robot, const double *action, const double *time_now,
                const double *consts, double *out) {
    for (int e = 0; e < E; ++e) {
        const double *rb = robot + (size_t)e * NS;
        const double *ac = action + 2 * e;
        double T = consts[5];
        double dmin = INFINITY; int collision = 0;
        for (int i = 0; i < n; ++i) {
            const double *h = states + ((size_t)e * rows + i) * NS;
            double dx = h[0] - rb[0], dy = h[1] - rb[1];
            double vx = h[3] - ac[0], vy = h[4] - ac[1];
            double ex = dx + vx * T, ey = dy + vy * T;
            double cd = point_to_segment_dist(dx, dy, ex, ey, 0, 0) - h[8] - rb[8];
            if (cd < 0) { collision = 1; break; }
            else if (cd < dmin) dmin = cd;
        }
        double endx = rb[0] + ac[0] * T, endy = rb[1] + ac[1] * T;
        int goal = np_norm(1, endx - rb[10], endy - rb[11]) < rb[8];
        double reward; int term, trunc, code;
        if (time_now[e] >= consts[0] - 1) { reward = 0; trunc = 1; term = 0; code = 1; }
        else if (collision) { reward = consts[1]; trunc = 0; term = 1; code = 2; }
        else if (goal) { reward = consts[2]; trunc = 0; term = 1; code = 3; }
        else if (dmin < consts[3]) { reward = (dmin - consts[3]) * consts[4] * T; trunc = 0; term = 0; code = 4; }
        else { reward = 0; trunc = 0; term = 0; code = 0; }
        double admin = 10000.0; int acol = 0, kcol = 0;
        for (int i = 0; i < n; ++i) {
            const double *h = states + ((size_t)e * rows + i) * NS;
            double nr = np_norm(1, h[0] - rb[0], h[1] - rb[1]);
            double d = nr - h[8] - rb[8];
            admin = d < admin ? d : admin;
            if (admin <= 0) acol = 1;
            if (nr < (h[8] + rb[8])) kcol = 1;
        }
        int agoal = np_norm(1, rb[0] - rb[10], rb[1] - rb[11]) < rb[8];
        double *o = out + 12 * (size_t)e;
        o[0] = collision; o[1] = dmin; o[2] = goal; o[3] = reward; o[4] = term; o[5] = trunc; o[6] = code;
        o[7] = acol; o[8] = admin; o[9] = agoal; o[10] = kcol; o[11] = 0;
    }
}

/* src/sensors.py:53-69 with :24-33 (circle) and :35-51 (segment). humans [E][n][3] = x,y,r;
 * walls [W][S][2][2] NaN padded (shared or per env); pose [E][3] = x,y,yaw. ranges/hits [E][samples].
 * hit index: humans 0..n-1, then wall segments in (wall, segment-slot) order counting only non-NaN slots,
 * offset by n; -1 when nothing is closer than max_distance (strict '<', first winner). */
void orc_laser(int E, int n, int W, int S, int walls_per_env, const double *humans, const double *walls, const double *pose,
               double range, int samples, double max_distance, double *ranges, int64_t *hits, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
    for (int e = 0; e < E; ++e) {
        const double *ps = pose + 3 * (size_t)e;
        const double *wl = walls + (walls_per_env ? (size_t)e * W * S * 4 : 0);
        double start = ps[2] - (range / 2), stop = ps[2] + (range / 2);
        int div = samples - 1;
        double step = div > 0 ? (stop - start) / div : 0.0; /* numpy.linspace: start + arange*step, last = stop */
        for (int k = 0; k < samples; ++k) {
            double ang = (k == div && div > 0) ? stop : start + k * step;
            double dxr = cos(ang), dyr = sin(ang);
            double m = max_distance; int64_t hit = -1;
            for (int i = 0; i < n; ++i) {
                const double *h = humans + ((size_t)e * n + i) * 3;
                double sx = ps[0] - h[0], sy = ps[1] - h[1];
                double b = np_dot(1, sx, sy, dxr, dyr);
                double cc = np_dot(1, sx, sy, sx, sy) - (h[2] * h[2]);
                double hh = b * b - cc;
                double rc;
                if (hh < 0.0) rc = max_distance;
                else { hh = sqrt(hh); double t = -b - hh; rc = t < 0.0 ? max_distance : (t < max_distance ? t : max_distance); }
                if (rc < m) { m = rc; hit = i; }
            }
            int64_t idx = n;
            for (int w = 0; w < W; ++w)
                for (int s = 0; s < S; ++s) {
                    const double *sg = wl + ((size_t)w * S + s) * 4;
                    if (isnan(sg[0])) continue;
                    double x1 = sg[0], y1 = sg[1], x2 = sg[2], y2 = sg[3];
                    double x3 = ps[0], y3 = ps[1], x4 = ps[0] + dxr, y4 = ps[1] + dyr;
                    double den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
                    double rc = max_distance;
                    if (!(den <= 0.0)) {
                        double t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / den;
                        double u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / den;
                        if (0 < t && t < 1 && u > 0) {
                            double ix = x1 + t * (x2 - x1), iy = y1 + t * (y2 - y1);
                            double d = np_norm(1, ps[0] - ix, ps[1] - iy);
                            rc = d < max_distance ? d : max_distance;
                        }
                    }
                    if (rc < m) { m = rc; hit = idx; }
                    ++idx;
                }
            ranges[(size_t)e * samples + k] = m;
            hits[(size_t)e * samples + k] = hit;
        }
    }
}

/* crowd_nav/policy/cadrl.py:42-83 compute_rotated_states_and_reward with :13-40 transform_state_to_agent_centric (Numba: plain
 * arithmetic).  cur [E][N][5|7] = px,py,vx,vy,r(,theta,omega); next [E][N][4|6] = x,y,vx,vy | x,y,yaw,vx,vy,omega;
 * robot [E][9] = px,py,vx,vy,r,gx,gy,vd,theta; actions [A][2]; rotated [E][A][N][13|15]; rewards [E][A]. */
void orc_lookahead(int E, int N, int A, int visible, const double *cur, const double *next, const double *robot, const double *actions,
                   double dt, double *rotated, double *rewards, int n_threads) {
    const int cw = visible ? 7 : 5, nw = visible ? 6 : 4, ow = visible ? 15 : 13;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
    for (int e = 0; e < E; ++e) {
        const double *rb = robot + 9 * (size_t)e;
        for (int a = 0; a < A; ++a) {
            const double ax = actions[2 * a], ay = actions[2 * a + 1];
            const double npx = rb[0] + ax * dt, npy = rb[1] + ay * dt;
            double dmin = (double)INT64_MAX; int collision = 0;
            for (int j = 0; j < N; ++j) {
                const double *h = cur + ((size_t)e * N + j) * cw;
                const double dx = h[0] - rb[0], dy = h[1] - rb[1];
                const double ex = dx + (h[2] - ax) * dt, ey = dy + (h[3] - ay) * dt;
                const double dist = point_to_segment_dist(dx, dy, ex, ey, 0, 0) - h[4] - rb[4];
                if (dist < 0) { collision = 1; break; }
                else if (dist >= 0 && dist < dmin) dmin = dist;
            }
            const int reached = norm2(npx - rb[5], npy - rb[6]) < rb[4];
            double rew;
            if (collision) rew = -0.25; else if (reached) rew = 1; else if (dmin < 0.2) rew = (dmin - 0.2) * 0.5 * dt; else rew = 0;
            rewards[(size_t)e * A + a] = rew;
            const double rot = atan2(rb[6] - npy, rb[5] - npx), c = cos(rot), s = sin(rot);
            for (int j = 0; j < N; ++j) {
                const double *h = cur + ((size_t)e * N + j) * cw;
                const double *nx = next + ((size_t)e * N + j) * nw;
                const double hx = nx[0], hy = nx[1], hvx = visible ? nx[3] : nx[2], hvy = visible ? nx[4] : nx[3];
                double *o = rotated + (((size_t)e * A + a) * N + j) * ow;
                o[0] = norm2(rb[5] - npx, rb[6] - npy); o[1] = rb[7]; o[2] = 0.0; o[3] = rb[4];
                o[4] = ax * c + ay * s; o[5] = ay * c - ax * s;
                o[6] = (hx - npx) * c + (hy - npy) * s; o[7] = (hy - npy) * c - (hx - npx) * s;
                o[8] = hvx * c + hvy * s; o[9] = hvy * c - hvx * s;
                o[10] = h[4]; o[11] = norm2(hx - npx, hy - npy); o[12] = rb[4] + h[4];
                if (visible) { o[13] = nx[2] - o[2]; o[14] = nx[5]; }
            }
        }
    }
}

/* src/robot_agent.py:35-48 RobotAgent.check_collisions: the robot is pushed out of every human it overlaps, in list
 * order, then out of every wall polygon it overlaps (closest point of obstacle.py:53-66), each push seeing the previous ones.
 * humans [E][n][3] = x,y,r; walls [W][S][4] NaN padded (shared by the envs); robot [E][3] = x,y,r, position updated in place. */
void orc_robot_push_out(int E, int n, int W, int S, const double *humans, const double *walls, double *robot) {
    for (int e = 0; e < E; ++e) {
        double *rb = robot + 3 * (size_t)e;
        for (int j = 0; j < n; ++j) {
            const double *h = humans + ((size_t)e * n + j) * 3;
            const double dx = rb[0] - h[0], dy = rb[1] - h[1];
            const double dist = np_norm(1, dx, dy);
            if (dist < h[2] + rb[2]) {
                const double ux = dx / dist, uy = dy / dist, sum = h[2] + rb[2];
                rb[0] = h[0] + ux * sum; rb[1] = h[1] + uy * sum;
            }
        }
        for (int w = 0; w < W; ++w) {
            double cx, cy;
            closest_point(walls + (size_t)w * S * 4, S, rb[0], rb[1], 0, &cx, &cy);
            int any = 0;
            for (int s = 0; s < S; ++s) any |= !isnan(walls[((size_t)w * S + s) * 4]);
            const double dist = any ? np_norm(1, cx - rb[0], cy - rb[1]) : 10000.0;   /* min_distance of obstacle.py:54,64 */
            if (dist < rb[2]) {
                const double nn = np_norm(1, cx - rb[0], cy - rb[1]);
                const double ux = (rb[0] - cx) / nn, uy = (rb[1] - cy) / nn;
                rb[0] = cx + ux * rb[2]; rb[1] = cy + uy * rb[2];
            }
        }
    }
}

int orc_abi_version(void) { return 1; }
