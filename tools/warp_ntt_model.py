"""Bit-exact model (plain Python, small prime field) of the lane / register-slot bookkeeping of csrc/ntt_warp.cuh: the same
swap schedule, twiddle indices and output-row formula, checked against the O(R^2) definition of the DFT for every (EL, K) the
library instantiates.  Run on the CPU: `python tools/warp_ntt_model.py` (also executed by tests/test_warp_ntt_model.py)."""
import itertools
import random

P = 65537          # 2^16 + 1: has 2^k-th roots of unity for k <= 16
G = 3


def slot_logical(EL, K, s, t):
    cur = K - EL + s
    for i in range(EL, t):
        if (i - EL) % EL == s:
            cur = K - 1 - i
    return cur


def lane_logical_final(EL, K, l):
    t = K - 1 - l
    return slot_logical(EL, K, (t - EL) % EL, t)


def out_row(EL, K, lane, slot):
    kr = 0
    for l in range(K - EL):
        kr |= ((lane >> l) & 1) << (K - 1 - lane_logical_final(EL, K, l))
    for s in range(EL):
        kr |= ((slot >> s) & 1) << (K - 1 - slot_logical(EL, K, s, K))
    return kr


def warp_ntt(EL, K, column):
    """column: list of R = 2^K values (natural order).  Returns the list of (output row, value) the kernel would store."""
    E, LB, R = 1 << EL, K - EL, 1 << K
    w_r = pow(G, (P - 1) // R, P)
    tw = [pow(w_r, j, P) for j in range(R // 2)]
    lanes = 1 << LB
    x = [[column[(s << LB) | l] for s in range(E)] for l in range(lanes)]       # x[lane][slot]
    for T in range(K):
        B = K - 1 - T
        SIG = (EL - 1 - T) if T < EL else (T - EL) % EL
        if T >= EL:
            new = [row[:] for row in x]
            for l in range(lanes):
                hi = (l >> B) & 1
                partner = l ^ (1 << B)
                for s0 in range(E):
                    if s0 & (1 << SIG):
                        continue
                    s1 = s0 | (1 << SIG)
                    p_hi = (partner >> B) & 1
                    recv = x[partner][s0] if p_hi else x[partner][s1]            # what the partner sends
                    if hi:
                        new[l][s0] = recv
                    else:
                        new[l][s1] = recv
            x = new
        for l in range(lanes):
            for s0 in range(E):
                if s0 & (1 << SIG):
                    continue
                s1 = s0 | (1 << SIG)
                if T >= EL:
                    low = l & ((1 << B) - 1)
                else:
                    low = ((s0 & ((1 << SIG) - 1)) << LB) | l
                a, b = x[l][s0], x[l][s1]
                x[l][s0] = (a + b) % P
                d = (a - b) % P
                if B > 0:
                    d = d * tw[low << (K - 1 - B)] % P
                x[l][s1] = d
    return [(out_row(EL, K, l, s), x[l][s]) for l in range(lanes) for s in range(E)]


def check(EL, K, seed=0):
    R = 1 << K
    rnd = random.Random(seed)
    col = [rnd.randrange(P) for _ in range(R)]
    w_r = pow(G, (P - 1) // R, P)
    want = [sum(col[j] * pow(w_r, i * j, P) for j in range(R)) % P for i in range(R)]
    got = [None] * R
    for row, v in warp_ntt(EL, K, col):
        assert got[row] is None, ("output row written twice", EL, K, row)
        got[row] = v
    assert got == want, (EL, K)


def main():
    for EL, K in itertools.product((1, 2, 3), range(2, 9)):
        if K > EL and K - EL <= 5:
            check(EL, K, seed=K * 10 + EL)
    print("warp NTT bookkeeping model: ok")


if __name__ == "__main__":
    main()
