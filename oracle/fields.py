"""Field / curve constants and domain helpers (oracle; test infrastructure only).

Follows /root/reference/python/zksnake/constant.py:5-15 for the four moduli and arkworks'
`Radix2EvaluationDomain::new` semantics (size = next power of two, group_gen =
TWO_ADIC_ROOT_OF_UNITY^(2^(s-log N)), reached from /root/reference/src/bn254/polynomial.rs:541).
"""

BN254, BLS12_381 = 0, 1
CURVE_NAMES = {BN254: "BN254", BLS12_381: "BLS12_381"}


class CurveParams:
    def __init__(self, cid, q, r, fr_gen, two_adicity, b_g1, g1, xi, b_g2, g2):
        self.id = cid
        self.q = q  # base field modulus
        self.r = r  # scalar field modulus
        self.fr_gen = fr_gen  # multiplicative generator of Fr* used by arkworks
        self.two_adicity = two_adicity
        self.b_g1 = b_g1  # y^2 = x^3 + b
        self.g1 = g1
        self.xi = xi  # Fq2 non-residue used for the twist (as (c0, c1))
        self.b_g2 = b_g2  # (c0, c1)
        self.g2 = g2  # ((x0, x1), (y0, y1))
        self.two_adic_root = pow(fr_gen, (r - 1) >> two_adicity, r)
        self.fq_bytes = (q.bit_length() + 7) // 8

    def omega(self, log_n):
        """group_gen of the radix-2 domain of size 2^log_n."""
        if log_n > self.two_adicity:
            raise ValueError("Domain size is too large")
        return pow(self.two_adic_root, 1 << (self.two_adicity - log_n), self.r)


_bn_q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_bn_r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_bls_q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_bls_r = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _fq2_mul(a, b, q):
    return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)


def _fq2_inv(a, q):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, q)
    return (a[0] * d % q, (-a[1]) * d % q)


PARAMS = {
    BN254: CurveParams(
        BN254, _bn_q, _bn_r, 5, 28, 3, (1, 2), (9, 1),
        _fq2_mul((3, 0), _fq2_inv((9, 1), _bn_q), _bn_q),
        ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
          11559732032986387107991004021392285783925812861821192530917403151452391805634),
         (8495653923123431417604973247489272438418190587263600148770280649306958101930,
          4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    ),
    BLS12_381: CurveParams(
        BLS12_381, _bls_q, _bls_r, 7, 32, 4,
        (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
         0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1),
        (1, 1), (4, 4),
        ((0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
          0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
         (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
          0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)),
    ),
}


def curve_id(name):
    """Same aliases as /root/reference/python/zksnake/ecc.py:12-37."""
    return {"BN128": BN254, "BN254": BN254, "ALT_BN128": BN254, "BLS12_381": BLS12_381}[name]


def next_power_of_two(n):
    """/root/reference/python/zksnake/utils.py:26-28 (quirks for n<=0 kept)."""
    return 1 << (n - 1).bit_length()


def domain_log(size):
    """log2 of the arkworks radix-2 domain for a requested size (size=0 -> domain of size 1)."""
    if size <= 1:
        return 0
    return (size - 1).bit_length()
