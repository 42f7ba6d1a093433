"""Function layer over the `_algebra` polynomial modules: the host-side counterpart of the reference's
python/zksnake/polynomial.py (Polynomial factory :17-58, fft/ifft/coset wrappers :90-123, _pad_coeffs/mul_over_fft :126-165,
add/mul_over_evaluation_domain :168-185, barycentric_eval :204-216).  Signatures take the field modulus `p` like the
reference's; transforms and pointwise products run in libzkb200.so."""
from ._algebra import polynomial_bls12_381, polynomial_bn254

BN254_SCALAR_FIELD = polynomial_bn254.MODULUS
BLS12_381_SCALAR_FIELD = polynomial_bls12_381.MODULUS
POLY_OBJECT = {BN254_SCALAR_FIELD: polynomial_bn254, BLS12_381_SCALAR_FIELD: polynomial_bls12_381}


def next_power_of_two(n):
    """utils.py:26-28, including its behaviour for n <= 0 (Python's bit_length of negatives): 0 -> 2, -1 -> 4."""
    return 1 << (n - 1).bit_length()


def Polynomial(coeffs, p, domain_size=None):
    """Dense univariate polynomial c[0] + c[1] x + ... over Z_p with an attached evaluation-domain size (default: len)."""
    if not isinstance(coeffs, list):
        raise TypeError("Coefficients must be in list or dict")   # multivariate dicts are outside the proving path
    if not domain_size:
        domain_size = len(coeffs)
    return POLY_OBJECT[p].Polynomial(1, coeffs, domain_size)   # (the mirror also takes bare ints: no per-coefficient tuple)


def get_evaluation_point(domain, i, p):
    return 1 if i == 0 else POLY_OBJECT[p].get_evaluation_point(domain, i)


def get_all_evaluation_points(domain, p):
    return POLY_OBJECT[p].get_all_evaluation_points(domain)


def fft(coeffs, p, size=None):
    return POLY_OBJECT[p].fft(coeffs, size or len(coeffs))


def coset_fft(coeffs, p, size=None):
    return POLY_OBJECT[p].coset_fft(coeffs, size or len(coeffs))


def ifft(evals, p, size=None):
    return POLY_OBJECT[p].ifft(evals, size or len(evals))


def coset_ifft(evals, p, size=None):
    return POLY_OBJECT[p].coset_ifft(evals, size or len(evals))


def _pad_coeffs(a, b):
    """polynomial.py:126-148: the padding rule that decides the FFT domain of mul_over_fft.  Reproduced, not "fixed":
    equal degrees d get next_power_of_two(d) zeros each (so two degree-0 operands get 2, empty lists get 4); otherwise the
    longer one gets next_power_of_two(max degree) zeros and the shorter is padded to the same length."""
    da, db = len(a) - 1, len(b) - 1
    if da == db:
        pad = next_power_of_two(da)
        return a + [0] * pad, b + [0] * pad
    length = next_power_of_two(max(da, db))
    if da > db:
        return a + [0] * length, b + [0] * (da + length - db)
    return a + [0] * (db + length - da), b + [0] * length


def mul_over_fft(domain, a, b, p, return_poly=True):
    """Product of two Polynomials through NTT -> pointwise -> iNTT (polynomial.py:151-165); the result keeps `domain`."""
    ca, cb = _pad_coeffs(a.coeffs(), b.coeffs())
    fa, fb = fft(ca, p), fft(cb, p)
    prod = mul_over_evaluation_domain(len(fa), fa, fb, p)
    return Polynomial(ifft(prod, p), p, domain) if return_poly else prod


def add_over_evaluation_domain(domain, evals, p):
    mod = POLY_OBJECT[p]
    acc = evals[0]
    for e in evals[1:]:
        acc = mod.add_over_evaluation_domain(domain, acc, e)
    return acc


def mul_over_evaluation_domain(domain, a, b, p):
    return POLY_OBJECT[p].mul_over_evaluation_domain(domain, a, b)


def evaluate_vanishing_polynomial(domain, x, p):
    return POLY_OBJECT[p].evaluate_vanishing_polynomial(domain, x)


def evaluate_lagrange_coefficients(domain, x, p):
    return POLY_OBJECT[p].evaluate_lagrange_coefficients(domain, x)


def barycentric_eval(domain, sparse_eval, x, p):
    """Value at x of the polynomial given by a few non-zero evaluations {index: value} on the size-`domain` subgroup."""
    omega = get_evaluation_point(domain, 1, p)
    total = 0
    for i, v in sparse_eval.items():
        w = pow(omega, i, p)
        total += v * w * pow(x - w, -1, p)
    return (pow(x, domain, p) - 1) * pow(domain, -1, p) * total % p
