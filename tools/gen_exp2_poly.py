"""Generate near-minimax polynomial coefficients for 2^r on |r| <= h (Remez exchange, relative error).
Usage: python tools/gen_exp2_poly.py DEGREE HALFWIDTH
Used to produce the constants in gingr_b200/csrc/exp2_poly.cuh."""
import sys
import mpmath as mp

mp.mp.dps = 60


def remez(f, deg, a, b, iters=30):
    n = deg + 2
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (n - 1 - i) / (n - 1)) for i in range(n)]
    coeffs = None
    for _ in range(iters):
        # solve sum c_k x^k + (-1)^i E f(x_i) = f(x_i)   (relative error equioscillation)
        A = mp.matrix(n, n)
        rhs = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for k in range(deg + 1):
                A[i, k] = x ** k
            A[i, deg + 1] = (-1) ** i * f(x)
            rhs[i] = f(x)
        sol = mp.lu_solve(A, rhs)
        coeffs = [sol[k] for k in range(deg + 1)]
        err = lambda x: (mp.polyval(coeffs[::-1], x) - f(x)) / f(x)
        # find extrema of err on a fine grid between sign changes
        grid = [a + (b - a) * i / 4000 for i in range(4001)]
        vals = [err(x) for x in grid]
        ext = []
        i0 = 0
        sign = mp.sign(vals[0]) or 1
        seg_best = 0
        for i, v in enumerate(vals):
            s = mp.sign(v) or sign
            if s != sign:
                ext.append(seg_best)
                sign = s
                seg_best = i
            elif abs(v) > abs(vals[seg_best]):
                seg_best = i
        ext.append(seg_best)
        if len(ext) != n:
            break
        new_xs = []
        for i in ext:
            lo = grid[max(i - 1, 0)]
            hi = grid[min(i + 1, 4000)]
            # golden refine on |err|
            try:
                x = mp.findroot(lambda t: mp.diff(err, t), grid[i]) if 0 < i < 4000 else grid[i]
                if not (lo <= x <= hi):
                    x = grid[i]
            except Exception:
                x = grid[i]
            new_xs.append(x)
        if max(abs(p - q) for p, q in zip(xs, new_xs)) < mp.mpf(10) ** -30:
            xs = new_xs
            break
        xs = new_xs
    maxerr = max(abs(err(a + (b - a) * i / 20000)) for i in range(20001))
    return coeffs, maxerr


if __name__ == "__main__":
    deg = int(sys.argv[1])
    h = mp.mpf(sys.argv[2])
    f = lambda x: mp.mpf(2) ** x
    c, e = remez(f, deg, -h, h)
    print(f"// 2^r, |r| <= {sys.argv[2]}, degree {deg}, max relative error {mp.nstr(e, 5)}")
    for k, ck in enumerate(c):
        print(f"  {mp.nstr(ck, 20)},  // c{k}  ({float(ck).hex()})")
