"""Dev tool: polynomial coefficients of the kernels' own fp64 atan / sin / cos (csrc/snp_math.cuh), by interpolation at Chebyshev
nodes in 60-digit arithmetic (within a small factor of minimax), rounded to double, with the error of the double-precision Horner
form measured on a dense grid.

    python tools/gen_coeffs.py
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(f, a, b, n):
    """Coefficients c_0..c_{n-1} (monomial basis in x) of the degree n-1 interpolant of f at the Chebyshev nodes of [a, b]."""
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[j] for j in range(n)]


def horner_np(c, w):
    p = np.full_like(w, float(c[-1]))
    for k in range(len(c) - 2, -1, -1):
        p = p * w + float(c[k])
    return p


def report(name, coeffs):
    print(f"// {name}")
    print("    " + ", ".join(f"{float(c)!r}" for c in coeffs))


def main():
    # atan(z) = z + z * w * Q(w), w = z^2 in [0, 1]
    def q_atan(w):
        if w == 0:
            return mp.mpf(-1) / 3
        z = mp.sqrt(w)
        return (mp.atan(z) / z - 1) / w
    for n in (16, 18, 19, 20, 22):
        c = cheb_fit(q_atan, mp.mpf(0), mp.mpf(1), n)
        z = np.linspace(1e-6, 1.0, 200001)
        w = z * z
        approx = z + z * (w * horner_np(c, w))
        ref = np.array([float(mp.atan(mp.mpf(float(t)))) for t in z[::97]])
        err = np.abs(approx[::97] - ref) / np.spacing(ref)
        # truncation error alone (high precision evaluation of the rounded coefficients)
        tr = max(abs(sum(mp.mpf(float(ck)) * mp.mpf(wk) ** k for k, ck in enumerate(c)) - q_atan(mp.mpf(wk))) for wk in np.linspace(0, 1, 401))
        print(f"atan n={n}: max ulp (double Horner, no fma) {err.max():.2f}, truncation of Q {float(tr):.2e}")
        if n == 20:
            report("atan Q, 20 coefficients", c)
    # sin(r) = r + r^3 S(z), cos(r) = 1 - z/2 + z^2 C(z), z = r^2, |r| <= pi/4
    lim = (mp.pi / 4) ** 2 * mp.mpf("1.02")

    def s_sin(z):
        if z == 0:
            return mp.mpf(-1) / 6
        r = mp.sqrt(z)
        return (mp.sin(r) - r) / (r * z)

    def c_cos(z):
        if z == 0:
            return mp.mpf(1) / 24
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z)
    for n in (6, 7):
        cs = cheb_fit(s_sin, mp.mpf(0), lim, n)
        cc = cheb_fit(c_cos, mp.mpf(0), lim, n)
        r = np.linspace(-0.7854, 0.7854, 100001)
        z = r * r
        s = r + r * z * horner_np(cs, z)
        c = 1.0 - 0.5 * z + z * z * horner_np(cc, z)
        rs = np.array([float(mp.sin(mp.mpf(float(t)))) for t in r[::53]])
        rc = np.array([float(mp.cos(mp.mpf(float(t)))) for t in r[::53]])
        print(f"sincos n={n}: sin max ulp {np.max(np.abs(s[::53] - rs) / np.spacing(np.abs(rs) + 1e-300)):.2f}, "
              f"cos max ulp {np.max(np.abs(c[::53] - rc) / np.spacing(rc)):.2f}")
        if n == 6:
            report("sin S, 6 coefficients", cs)
            report("cos C, 6 coefficients", cc)


if __name__ == "__main__":
    main()
