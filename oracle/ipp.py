"""Oracle: inner-product argument, restated from /root/reference/src/ipp.rs (test infrastructure).

Line-by-line restatement of `IPP::create_ipp` (ipp.rs:35-202), `IPP::verify_ipp`
(ipp.rs:204-260) and `IPP::verification_scalars` (ipp.rs:262-315) over the
oracle's integer arithmetic.  Pure function of its inputs (create_ipp samples no
randomness), so its outputs serve as known answers for the CUDA path.
"""


class VerificationError(Exception):
    pass


class IPPProof:
    def __init__(self, L, R, a, b):
        self.L, self.R, self.a, self.b = L, R, a, b


def create_ipp(C, transcript, Q, G_factors, H_factors, G_vec, H_vec, a_vec, b_vec):
    """ipp.rs:35-202."""
    n = len(G_vec)
    assert n & (n - 1) == 0 and n > 0                      # ipp.rs:48
    assert len(H_vec) == len(a_vec) == len(b_vec) == len(G_factors) == len(H_factors) == n  # ipp.rs:51-55
    r = C.r
    G, H, a, b = list(G_vec), list(H_vec), [x % r for x in a_vec], [x % r for x in b_vec]
    transcript.innerproduct_domain_sep(n)                  # ipp.rs:62
    L_vec, R_vec = [], []
    first = True
    while n != 1:                                          # ipp.rs:68 (first round) and ipp.rs:138
        n //= 2
        a_L, a_R = a[:n], a[n:]
        b_L, b_R = b[:n], b[n:]
        G_L, G_R = G[:n], G[n:]
        H_L, H_R = H[:n], H[n:]
        c_L = C.inner_product(a_L, b_R)                    # ipp.rs:77 / 145
        c_R = C.inner_product(a_R, b_L)                    # ipp.rs:78 / 146
        if first:
            Gf_L, Gf_R = G_factors[:n], G_factors[n:]
            Hf_L, Hf_R = H_factors[:n], H_factors[n:]
            L_s = [x * y % r for x, y in zip(a_L, Gf_R)] + [x * y % r for x, y in zip(b_R, Hf_L)] + [c_L]  # ipp.rs:80-83
            R_s = [x * y % r for x, y in zip(a_R, Gf_L)] + [x * y % r for x, y in zip(b_L, Hf_R)] + [c_R]  # ipp.rs:93-96
        else:
            L_s = a_L + b_R + [c_L]                        # ipp.rs:152-155
            R_s = a_R + b_L + [c_R]                        # ipp.rs:164-167
        L = C.msm(G_R + H_L + [Q], L_s)                    # ipp.rs:91 / 158
        R = C.msm(G_L + H_R + [Q], R_s)                    # ipp.rs:104 / 170
        transcript.commit_point(b"L", L)                   # ipp.rs:106 / 172
        transcript.commit_point(b"R", R)
        L_vec.append(L)
        R_vec.append(R)
        u = transcript.challenge_scalar(b"u")              # ipp.rs:112 / 178
        u_inv = C.fr_inv(u)
        for i in range(n):                                 # ipp.rs:115-130 / 181-188
            a_L[i] = (a_L[i] * u + u_inv * a_R[i]) % r
            b_L[i] = (b_L[i] * u_inv + u * b_R[i]) % r
            if first:
                G_L[i] = C.binary_scalar_mul(G_L[i], G_R[i], u_inv * Gf_L[i], u * Gf_R[i])
                H_L[i] = C.binary_scalar_mul(H_L[i], H_R[i], u * Hf_L[i], u_inv * Hf_R[i])
            else:
                G_L[i] = C.binary_scalar_mul(G_L[i], G_R[i], u_inv, u)
                H_L[i] = C.binary_scalar_mul(H_L[i], H_R[i], u, u_inv)
        a, b, G, H = a_L, b_L, G_L, H_L
        first = False
    return IPPProof(L_vec, R_vec, a[0], b[0])


def verification_scalars(C, L_vec, R_vec, n, transcript):
    """ipp.rs:262-315 -> (u_sq, u_inv_sq, s)."""
    lg_n = len(L_vec)
    if lg_n >= 32:                                         # ipp.rs:269-273
        raise VerificationError()
    if n != (1 << lg_n):                                   # ipp.rs:274-276
        raise VerificationError()
    transcript.innerproduct_domain_sep(n)
    challenges = []
    for L, R in zip(L_vec, R_vec):                         # ipp.rs:283-288
        transcript.commit_point(b"L", L)
        transcript.commit_point(b"R", R)
        challenges.append(transcript.challenge_scalar(b"u"))
    r = C.r
    inv, prod_inv = C.fr_batch_invert(challenges)          # ipp.rs:295
    sq = [u * u % r for u in challenges]
    inv_sq = [u * u % r for u in inv]
    s = [prod_inv]                                         # ipp.rs:303-312
    for i in range(1, n):
        lg_i = i.bit_length() - 1
        k = 1 << lg_i
        s.append(s[i - k] * sq[(lg_n - 1) - lg_i] % r)
    return sq, inv_sq, s


def verify_ipp(C, n, transcript, G_factors, H_factors, P, Q, G, H, a, b, L_vec, R_vec):
    """ipp.rs:204-260; raises VerificationError on failure."""
    u_sq, u_inv_sq, s = verification_scalars(C, L_vec, R_vec, n, transcript)
    r = C.r
    g_s = [(a * s_i) % r * g_i % r for g_i, s_i in zip(G_factors, s)][:len(G)]      # ipp.rs:220-224
    h_s = [(b * s_inv) % r * h_i % r for h_i, s_inv in zip(H_factors, reversed(s))]  # ipp.rs:227-232
    scalars = [a * b % r] + g_s + h_s + [(-x) % r for x in u_sq] + [(-x) % r for x in u_inv_sq]
    points = [Q] + list(G) + list(H) + list(L_vec) + list(R_vec)
    expected = C.msm(points, scalars)                      # ipp.rs:251-253
    if not C.eq(expected, P):                              # ipp.rs:255
        raise VerificationError()
