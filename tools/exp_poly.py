"""Coefficients of exp_pairwise's polynomial (rom-comma_b200/csrc/common.cuh): the degree-d Chebyshev interpolant of exp on
[-a, a], a = ln2/2 + margin, computed in 80-bit long double and rounded to double; prints the maximum relative error of the rounded
polynomial (evaluated in long double and by a double Horner scheme) next to the degree-11 Taylor polynomial."""
import math, sys
import numpy as np
ld = np.longdouble
a = ld('0.34662')


def fit(deg):
    k = np.arange(deg + 1, dtype=ld)
    x = np.cos((2 * k + 1) * ld(np.pi) / (2 * (deg + 1)))
    y = np.exp(x * a)
    T = np.zeros((deg + 1, deg + 1), dtype=ld)
    T[:, 0], T[:, 1] = 1, x
    for j in range(2, deg + 1):
        T[:, j] = 2 * x * T[:, j - 1] - T[:, j - 2]
    c = np.array([(2 if j else 1) * np.sum(y * T[:, j]) / (deg + 1) for j in range(deg + 1)], dtype=ld)
    polys = [np.array([1], dtype=ld), np.array([0, 1], dtype=ld)]
    for j in range(2, deg + 1):
        pj = np.zeros(j + 1, dtype=ld)
        pj[1:] += 2 * polys[j - 1]
        pj[:j - 1] -= polys[j - 2]
        polys.append(pj)
    mono = np.zeros(deg + 1, dtype=ld)
    for j in range(deg + 1):
        mono[:j + 1] += c[j] * polys[j]
    return (mono / np.array([a ** i for i in range(deg + 1)], dtype=ld)).astype(np.float64)


def max_rel_err(coef, r):
    rl = r.astype(ld)
    p = np.full_like(rl, ld(coef[-1]))
    for c in coef[-2::-1]:
        p = p * rl + ld(c)
    return float(np.max(np.abs(p / np.exp(rl) - 1)))


r = np.linspace(-0.34658, 0.34658, 200001)
for deg in [int(v) for v in sys.argv[1:] if v != 'table'] or ([] if 'table' in sys.argv[1:] else [9, 10]):
    m = fit(deg)
    print(deg, 'max relative error', max_rel_err(m, r))
    print('  ', ', '.join(f'{v:.17e}' for v in m))
if 'table' not in sys.argv[1:]:
    print('taylor 11', max_rel_err(np.array([1 / math.factorial(i) for i in range(12)]), r))


def fit_table_form(bits=5, deg=3):
    """exp_tab (common.cuh): x = (2^bits e + j) ln2/2^bits + r, exp(x) = 2^e T[j] (1 + r (1 + r g(r))): the degree-`deg` Chebyshev interpolant of
    g(r) = (e^r - 1 - r) / r^2 on |r| <= ln2 / 2^(bits+1) (2 per mille margin), coefficients rounded to double, and the table T[j] = 2^(j / 2^bits)."""
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    half = ld(math.log(2)) / (1 << (bits + 1)) * ld('1.002')

    def g(r):
        s, term = np.zeros_like(r), np.ones_like(r) / 2
        for k in range(2, 30):
            s = s + term
            term = term * r / (k + 1)
        return s

    k = np.arange(deg + 1, dtype=ld)
    x = np.cos((2 * k + 1) * ld(np.pi) / (2 * (deg + 1)))
    y = g(x * half)
    T = np.zeros((deg + 1, deg + 1), dtype=ld)
    T[:, 0] = 1
    T[:, 1] = x
    for j in range(2, deg + 1):
        T[:, j] = 2 * x * T[:, j - 1] - T[:, j - 2]
    c = np.array([(2 if j else 1) * np.sum(y * T[:, j]) / (deg + 1) for j in range(deg + 1)], dtype=ld)
    polys = [np.array([1], dtype=ld), np.array([0, 1], dtype=ld)]
    for j in range(2, deg + 1):
        pj = np.zeros(j + 1, dtype=ld)
        pj[1:] += 2 * polys[j - 1]
        pj[:j - 1] -= polys[j - 2]
        polys.append(pj)
    mono = np.zeros(deg + 1, dtype=ld)
    for j in range(deg + 1):
        mono[:j + 1] += c[j] * polys[j]
    coef = (mono / np.array([half ** i for i in range(deg + 1)], dtype=ld)).astype(np.float64)
    table = [float(Decimal(2) ** (Decimal(j) / Decimal(1 << bits))) for j in range(1 << bits)]
    return coef, table


def table_form_error(coef, bits=5):
    """max relative error of 1 + r (1 + r g(r)) against exp(r) on the reduced range (long double evaluation of the double coefficients)."""
    a = math.log(2) / (1 << (bits + 1))
    r = np.linspace(-a, a, 200001).astype(ld)
    s = np.full_like(r, ld(coef[-1]))
    for c in coef[-2::-1]:
        s = s * r + ld(c)
    return float(np.max(np.abs((1 + r * (1 + r * s)) / np.exp(r) - 1)))


if 'table' in sys.argv[1:]:
    coef, table = fit_table_form()
    print('table form: g coefficients c0..c3', ', '.join(f'{v:.17e}' for v in coef), 'max relative error', table_form_error(coef))
