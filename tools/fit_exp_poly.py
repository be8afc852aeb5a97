"""Coefficients of gperks_b200/csrc/gpx_math.cuh: degree-9 interpolant q of (e^r - 1 - r)/r^2 at the Chebyshev nodes of
|r| <= 1.001 ln2/2 (exp(r) = 1 + r + r^2 q(r)), rounded to double, and the ln2 split.  Needs mpmath."""
import mpmath as mp, struct
mp.mp.dps = 60
a = mp.log(2)/2 * mp.mpf('1.001')
def f(r):
    if r == 0: return mp.mpf(1)/2
    return (mp.exp(r) - 1 - r)/(r*r)
n = 9
nodes = [a*mp.cos(mp.pi*(2*k+1)/(2*(n+1))) for k in range(n+1)]
A = mp.matrix(n+1, n+1); b = mp.matrix(n+1, 1)
for i, x in enumerate(nodes):
    for j in range(n+1): A[i, j] = x**j
    b[i] = f(x)
c = mp.lu_solve(A, b)
cd = [float(c[j]) for j in range(n+1)]
for j in range(n+1):
    print(j+2, repr(cd[j]), hex(struct.unpack('<Q', struct.pack('<d', cd[j]))[0]), mp.nstr(c[j]*mp.factorial(j+2), 20))
mx = 0
for k in range(8001):
    x = -mp.log(2)/2 + mp.log(2)*k/8000
    p = 1 + x + x*x*sum(mp.mpf(cd[j])*x**j for j in range(n+1))
    mx = max(mx, abs(p/mp.exp(x) - 1))
print("max rel err (rounded coeffs, exact eval):", mp.nstr(mx, 5))
# ln2 split: hi with 32 trailing zero bits so n*hi exact for |n| < 2^11
ln2 = mp.log(2)
hi = float(ln2); hb = struct.unpack('<Q', struct.pack('<d', hi))[0] & ~((1<<27)-1); hi = struct.unpack('<d', struct.pack('<Q', hb))[0]
lo = float(ln2 - mp.mpf(hi))
print("ln2_hi", repr(hi), "ln2_lo", repr(lo), "log2e", repr(float(1/ln2)))
