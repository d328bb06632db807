"""Derives the series coefficients kLg / kEx of v2r_b200/csrc/idw_math.cuh and their design errors (mpmath, 200 bits):
interpolation at the Chebyshev nodes of the reduced argument's interval, which is within a few percent of minimax.

    log2(1+r)/r   on |r| <= 2^-(LB+1) (1 + 2^-20)   ->  D coefficients, error quoted in log2 units (times |r|)
    (2^(g/2^EB) - 1)/g  on |g| <= 1/2                ->  D coefficients, error quoted relative to the result
The shipped choice is LB = 10, D = 3 (5.1e-15) and EB = 9, D = 3 (2.2e-15); the table prints the neighbours that were weighed.
"""
import mpmath as mp

mp.mp.prec = 200


def cheb_fit(f, R, D):
    xs = [R * mp.cos(mp.pi * (2 * k + 1) / (2 * D)) for k in range(D)]
    A, b = mp.matrix(D, D), mp.matrix(D, 1)
    for i, x in enumerate(xs):
        for j in range(D):
            A[i, j] = x ** j
        b[i] = f(x)
    c = mp.lu_solve(A, b)
    return [c[j] for j in range(D)]


def maxerr(f, c, R, scale):
    m = 0
    for k in range(-2001, 2002, 2):
        x = R * k / 2001
        p = sum(cj * x ** j for j, cj in enumerate(c))
        m = max(m, abs((p - f(x)) * scale(x)))
    return m


def flog(r):
    return mp.log(1 + r, 2) / r if abs(r) > mp.mpf(10) ** -40 else 1 / mp.log(2)


for LB, D in [(7, 5), (8, 4), (9, 4), (10, 3), (10, 4), (11, 3)]:
    R = mp.mpf(2) ** -(LB + 1) * (1 + mp.mpf(2) ** -20)
    c = cheb_fit(flog, R, D)
    print(f"log2: table 2^{LB}, {D} coefficients: abs error {mp.nstr(maxerr(flog, c, R, abs), 3)}")
    print("   {" + ", ".join(mp.nstr(x, 20) for x in c) + "}")
for EB, D in [(5, 5), (8, 3), (8, 4), (9, 3), (10, 3)]:
    R = mp.mpf(0.5) * (1 + mp.mpf(2) ** -20)

    def fexp(g, EB=EB):
        return (mp.mpf(2) ** (g / 2 ** EB) - 1) / g if abs(g) > mp.mpf(10) ** -40 else mp.log(2) / 2 ** EB
    c = cheb_fit(fexp, R, D)
    print(f"exp2: table 2^{EB}, {D} coefficients: rel error {mp.nstr(maxerr(fexp, c, R, abs), 3)}")
    print("   {" + ", ".join(mp.nstr(x, 20) for x in c) + "}")
