"""Oracle: R1CS prover / verifier and the range gadgets (test infrastructure).

Restates `/root/reference/src/r1cs/prover.rs:84-129,142-184,300-593,604-700`,
`src/r1cs/verifier.rs:97-132,149-193,245-457`, `src/utils/vector_poly.rs:65-119`,
`src/r1cs/linear_combination.rs`, and the workload gadgets
`src/r1cs/gadgets/helper_constraints/positive_no.rs:8-40`,
`src/r1cs/gadgets/bound_check.rs:13-178`.

The reference samples blinding scalars with `FieldElement::random()` (OS entropy);
for reproducible parity every random draw is taken, in the reference's order
(SURVEY.md 8b "Randomness"), from an injected `rng()` callable.
"""
from . import ipp as ipp_mod

# Variable = (kind, index); kinds follow linear_combination.rs:11-23
COMMITTED, LEFT, RIGHT, OUTPUT, ONE = "V", "L", "R", "O", "1"
VAR_ONE = (ONE, 0)


class R1CSError(Exception):
    pass


class InvalidGeneratorsLength(R1CSError):
    pass


class MissingAssignment(R1CSError):
    pass


class LC:
    """LinearCombination: an ordered list of (Variable, coeff) terms (linear_combination.rs:35-37).
    Term order matters only for reproducing flattened_constraints exactly (it does not: sums commute),
    but it is kept identical to the reference anyway."""

    def __init__(self, terms=None):
        self.terms = list(terms or [])

    @staticmethod
    def of(x, C):
        if isinstance(x, LC):
            return x
        if isinstance(x, tuple):
            return LC([(x, 1)])                            # From<Variable>, linear_combination.rs:74-81
        return LC([(VAR_ONE, x % C.r)])                    # From<FieldElement>, :83-89

    def add(self, rhs, C):                                 # :152-159
        return LC(self.terms + LC.of(rhs, C).terms)

    def sub(self, rhs, C):                                 # :161-173
        return LC(self.terms + [(v, (-c) % C.r) for v, c in LC.of(rhs, C).terms])


class AllocatedQuantity:
    def __init__(self, variable, assignment):
        self.variable, self.assignment = variable, assignment


class R1CSProof:
    """proof.rs:26-58."""
    POINTS = ["A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2", "T_1", "T_3", "T_4", "T_5", "T_6"]
    SCALARS = ["t_x", "t_x_blinding", "e_blinding"]

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to_bytes(self, C):
        """Flat concatenation of to_bytes() of every element, in proof.rs field order, then the
        IPP proof's L[], R[], a, b.  (The reference derives serde only; this byte form is the
        fixture format of this repo.)"""
        out = b"".join(C.g1_to_bytes(getattr(self, k)) for k in self.POINTS)
        out += b"".join(C.fr_to_bytes(getattr(self, k)) for k in self.SCALARS)
        ip = self.ipp_proof
        out += b"".join(C.g1_to_bytes(p) for p in ip.L) + b"".join(C.g1_to_bytes(p) for p in ip.R)
        return out + C.fr_to_bytes(ip.a) + C.fr_to_bytes(ip.b)


    @classmethod
    def from_bytes(cls, C, buf):
        """inverse of to_bytes (fixture format of this repo)"""
        PB, SB = 1 + 2 * C.MODBYTES, C.MODBYTES
        lg = (len(buf) - 11 * PB - 5 * SB) // (2 * PB)
        assert len(buf) == 11 * PB + 5 * SB + 2 * lg * PB
        kw, off = {}, 0
        for k in cls.POINTS:
            kw[k] = C.g1_from_bytes(buf[off:off + PB])
            off += PB
        for k in cls.SCALARS:
            kw[k] = int.from_bytes(buf[off:off + SB], "big")
            off += SB
        L = [C.g1_from_bytes(buf[off + k * PB:off + (k + 1) * PB]) for k in range(lg)]
        off += lg * PB
        R = [C.g1_from_bytes(buf[off + k * PB:off + (k + 1) * PB]) for k in range(lg)]
        off += lg * PB
        a, b = int.from_bytes(buf[off:off + SB], "big"), int.from_bytes(buf[off + SB:off + 2 * SB], "big")
        return cls(ipp_proof=ipp_mod.IPPProof(L, R, a, b), **kw)


class Prover:
    def __init__(self, C, g, h, transcript):               # prover.rs:84-101
        transcript.r1cs_domain_sep()
        self.C, self.g, self.h, self.transcript = C, g, h, transcript
        self.constraints = []
        self.a_L, self.a_R, self.a_O = [], [], []
        self.v, self.v_blinding = [], []
        self.pending_multiplier = None
        self.deferred = []                                 # prover.rs:665-671 deferred_constraints

    def specify_randomized_constraints(self, callback):    # prover.rs:665-671 (1st phase) / :740-745 (inside a callback)
        if getattr(self, "randomizing", False):
            callback(self)
        else:
            self.deferred.append(callback)

    def challenge_scalar(self, label):                     # RandomizedConstraintSystem, prover.rs:759-763
        assert getattr(self, "randomizing", False)
        return self.transcript.challenge_scalar(label)

    def commit(self, v, v_blinding):                       # prover.rs:119-129
        C = self.C
        i = len(self.v)
        V = C.binary_scalar_mul(self.g, self.h, v, v_blinding)   # commit_to_field_element(g,h,v,r)
        self.v.append(v % C.r)
        self.v_blinding.append(v_blinding % C.r)
        self.transcript.commit_point(b"V", V)
        return V, (COMMITTED, i)

    # ---- ConstraintSystem (prover.rs:604-700)
    def eval(self, lc):                                    # prover.rs:282-296
        tot = 0
        for (k, i), c in lc.terms:
            val = {LEFT: self.a_L, RIGHT: self.a_R, OUTPUT: self.a_O, COMMITTED: self.v}[k][i] if k != ONE else 1
            tot += c * val
        return tot % self.C.r

    def _allocate_vars(self, l, r, o):
        i = len(self.a_L)
        self.a_L.append(l % self.C.r)
        self.a_R.append(r % self.C.r)
        self.a_O.append(o % self.C.r)
        return (LEFT, i), (RIGHT, i), (OUTPUT, i)

    def allocate_multiplier(self, input_assignments):      # prover.rs:651-659
        if input_assignments is None:
            raise MissingAssignment()
        l, r = input_assignments
        return self._allocate_vars(l, r, l * r)

    def multiply(self, left, right):                       # prover.rs:607-626
        l, r = self.eval(left), self.eval(right)
        lv, rv, ov = self._allocate_vars(l, r, l * r)
        self.constrain(LC(left.terms + [(lv, self.C.r - 1)]))
        self.constrain(LC(right.terms + [(rv, self.C.r - 1)]))
        return lv, rv, ov

    def constrain(self, lc):                               # prover.rs:661-665
        self.constraints.append(lc)

    def flattened_constraints(self, z):                    # prover.rs:142-184
        r = self.C.r
        n, m = len(self.a_L), len(self.v)
        wL, wR, wO, wV = [0] * n, [0] * n, [0] * n, [0] * m
        exp_z = z
        for lc in self.constraints:
            for (k, i), c in lc.terms:
                if k == LEFT:
                    wL[i] = (wL[i] + exp_z * c) % r
                elif k == RIGHT:
                    wR[i] = (wR[i] + exp_z * c) % r
                elif k == OUTPUT:
                    wO[i] = (wO[i] + exp_z * c) % r
                elif k == COMMITTED:
                    wV[i] = (wV[i] - exp_z * c) % r
            exp_z = exp_z * z % r
        return wL, wR, wO, wV

    def prove(self, G, H, rng):                            # prover.rs:322-593 (1- and 2-phase circuits)
        C, r, T = self.C, self.C.r, self.transcript
        T.append_u64(b"m", len(self.v))                    # :327
        n1 = len(self.a_L)
        if len(G) < n1:
            raise InvalidGeneratorsLength()
        i_blinding1, o_blinding1, s_blinding1 = rng(), rng(), rng()   # :336-338
        s_L1 = [rng() for _ in range(n1)]                  # :340
        s_R1 = [rng() for _ in range(n1)]                  # :341
        G1_, H1_ = G[:n1], H[:n1]
        # :347-355 commit_to_field_element_vectors(G,H,h,a_L,a_R,i_blinding)
        A_I1 = C.msm(G1_ + H1_ + [self.h], self.a_L + self.a_R + [i_blinding1])
        A_O1 = C.msm(G1_ + [self.h], self.a_O + [o_blinding1])         # :358
        S1 = C.msm(G1_ + H1_ + [self.h], s_L1 + s_R1 + [s_blinding1])  # :361-362
        T.commit_point(b"A_I1", A_I1)
        T.commit_point(b"A_O1", A_O1)
        T.commit_point(b"S1", S1)
        if not self.deferred:                              # create_randomized_constraints, :300-319
            T.r1cs_1phase_domain_sep()
        else:
            T.r1cs_2phase_domain_sep()
            callbacks, self.deferred = self.deferred, []
            self.randomizing = True
            for cb in callbacks:
                cb(self)
            self.randomizing = False
        n = len(self.a_L)
        n2 = n - n1
        padded_n = 1 if n == 0 else 1 << (n - 1).bit_length()   # usize::next_power_of_two
        pad = padded_n - n
        if len(G) < padded_n:
            raise InvalidGeneratorsLength()
        if n2 > 0:                                         # :384-427
            i_blinding2, o_blinding2, s_blinding2 = rng(), rng(), rng()
        else:
            i_blinding2 = o_blinding2 = s_blinding2 = 0    # :398-404
        s_L2 = [rng() for _ in range(n2)]                  # :401
        s_R2 = [rng() for _ in range(n2)]                  # :402
        if n2 > 0:
            G2_, H2_ = G[n1:n], H[n1:n]
            A_I2 = C.msm(G2_ + H2_ + [self.h], self.a_L[n1:] + self.a_R[n1:] + [i_blinding2])
            A_O2 = C.msm(G2_ + [self.h], self.a_O[n1:] + [o_blinding2])
            S2 = C.msm(G2_ + H2_ + [self.h], s_L2 + s_R2 + [s_blinding2])
        else:
            A_I2 = A_O2 = S2 = C.INF                       # :429
        s_L1, s_R1 = s_L1 + s_L2, s_R1 + s_R2              # s_L, s_R over all n multipliers (:474, :484)
        T.commit_point(b"A_I2", A_I2)
        T.commit_point(b"A_O2", A_O2)
        T.commit_point(b"S2", S2)
        y = T.challenge_scalar(b"y")                       # :438
        z = T.challenge_scalar(b"z")
        wL, wR, wO, wV = self.flattened_constraints(z)
        y_inv = C.fr_inv(y)
        exp_y_inv = C.vandermonde(y_inv, padded_n)         # :463
        l1, l2, l3 = [0] * n, [0] * n, [0] * n
        r0, r1, r3 = [0] * n, [0] * n, [0] * n
        exp_y = 1
        for i in range(n):                                 # :469-486
            l1[i] = (self.a_L[i] + exp_y_inv[i] * wR[i]) % r
            l2[i] = self.a_O[i]
            l3[i] = s_L1[i]
            r0[i] = (wO[i] - exp_y) % r
            r1[i] = (exp_y * self.a_R[i] + wL[i]) % r
            r3[i] = exp_y * s_R1[i] % r
            exp_y = exp_y * y % r
        ip = C.inner_product                               # vector_poly.rs:79-97
        t1 = ip(l1, r0)
        t2 = (ip(l1, r1) + ip(l2, r0)) % r
        t3 = (ip(l2, r1) + ip(l3, r0)) % r
        t4 = (ip(l1, r3) + ip(l3, r1)) % r
        t5 = ip(l2, r3)
        t6 = ip(l3, r3)
        t_1_b, t_3_b, t_4_b, t_5_b, t_6_b = rng(), rng(), rng(), rng(), rng()   # :490-494
        cm = lambda v, b: C.binary_scalar_mul(self.g, self.h, v, b)
        T_1, T_3, T_4, T_5, T_6 = cm(t1, t_1_b), cm(t3, t_3_b), cm(t4, t_4_b), cm(t5, t_5_b), cm(t6, t_6_b)
        for lab, P in ((b"T_1", T_1), (b"T_3", T_3), (b"T_4", T_4), (b"T_5", T_5), (b"T_6", T_6)):
            T.commit_point(lab, P)
        u = T.challenge_scalar(b"u")                       # :508
        x = T.challenge_scalar(b"x")
        t_2_b = ip(wV, self.v_blinding)                    # :513
        poly6 = lambda c: x * (c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * (c[4] + x * c[5]))))) % r  # vector_poly.rs:115-119
        t_x = poly6([t1, t2, t3, t4, t5, t6])
        t_x_blinding = poly6([t_1_b, t_2_b, t_3_b, t_4_b, t_5_b, t_6_b])
        l_vec = [(x * (l1[i] + x * (l2[i] + x * l3[i]))) % r for i in range(n)] + [0] * pad     # :526-527
        r_vec = [(r0[i] + x * (r1[i] + x * (x * r3[i]))) % r for i in range(n)]                 # :529
        for _ in range(n, padded_n):                       # :532-535
            r_vec.append((-exp_y) % r)
            exp_y = exp_y * y % r
        i_b = (i_blinding1 + u * i_blinding2) % r
        o_b = (o_blinding1 + u * o_blinding2) % r
        s_b = (s_blinding1 + u * s_blinding2) % r
        e_blinding = x * (i_b + x * (o_b + x * s_b)) % r   # :541
        T.commit_scalar(b"t_x", t_x)
        T.commit_scalar(b"t_x_blinding", t_x_blinding)
        T.commit_scalar(b"e_blinding", e_blinding)
        w = T.challenge_scalar(b"w")                       # :549
        Q = C.mul(self.g, w)
        G_factors = [1] * n1 + [u] * (n2 + pad)            # :552-556
        H_factors = [yi * g % r for yi, g in zip(exp_y_inv, G_factors)]   # :557-563
        if getattr(self, "trace", None) is not None:       # test hook: the intermediate vectors of steps 8, 12, 15 (row a11)
            self.trace.update(y=y, z=z, u=u, x=x, w=w, n=n, n1=n1, padded_n=padded_n, wL=wL, wR=wR, wO=wO, wV=wV, s_L=s_L1, s_R=s_R1,
                              l1=l1, l2=l2, l3=l3, r0=r0, r1=r1, r3=r3, t=[t1, t2, t3, t4, t5, t6], l_vec=l_vec, r_vec=r_vec,
                              G_factors=G_factors, H_factors=H_factors)
        ipp_proof = ipp_mod.create_ipp(C, T, Q, G_factors, H_factors, G[:padded_n], H[:padded_n], l_vec, r_vec)
        return R1CSProof(A_I1=A_I1, A_O1=A_O1, S1=S1, A_I2=A_I2, A_O2=A_O2, S2=S2, T_1=T_1, T_3=T_3, T_4=T_4,
                         T_5=T_5, T_6=T_6, t_x=t_x, t_x_blinding=t_x_blinding, e_blinding=e_blinding,
                         ipp_proof=ipp_proof)


class Verifier:
    def __init__(self, C, transcript):                     # verifier.rs:97-108
        transcript.r1cs_domain_sep()
        self.C, self.transcript = C, transcript
        self.num_vars = 0
        self.V = []
        self.constraints = []
        self.deferred = []                                 # verifier.rs:508-514

    def specify_randomized_constraints(self, callback):    # verifier.rs:508-514 / :577-582
        if getattr(self, "randomizing", False):
            callback(self)
        else:
            self.deferred.append(callback)

    def challenge_scalar(self, label):                     # verifier.rs:596-600
        assert getattr(self, "randomizing", False)
        return self.transcript.challenge_scalar(label)

    def commit(self, commitment):                          # verifier.rs:124-132
        i = len(self.V)
        self.transcript.commit_point(b"V", commitment)
        self.V.append(commitment)
        return (COMMITTED, i)

    def allocate_multiplier(self, _assignments):           # verifier.rs:538-548
        i = self.num_vars
        self.num_vars += 1
        return (LEFT, i), (RIGHT, i), (OUTPUT, i)

    def multiply(self, left, right):                       # verifier.rs:468-484
        lv, rv, ov = self.allocate_multiplier(None)
        self.constrain(LC(left.terms + [(lv, self.C.r - 1)]))
        self.constrain(LC(right.terms + [(rv, self.C.r - 1)]))
        return lv, rv, ov

    def constrain(self, lc):
        self.constraints.append(lc)

    def flattened_constraints(self, z):                    # verifier.rs:149-193
        r = self.C.r
        n, m = self.num_vars, len(self.V)
        wL, wR, wO, wV, wc = [0] * n, [0] * n, [0] * n, [0] * m, 0
        exp_z = z
        for lc in self.constraints:
            for (k, i), c in lc.terms:
                if k == LEFT:
                    wL[i] = (wL[i] + exp_z * c) % r
                elif k == RIGHT:
                    wR[i] = (wR[i] + exp_z * c) % r
                elif k == OUTPUT:
                    wO[i] = (wO[i] + exp_z * c) % r
                elif k == COMMITTED:
                    wV[i] = (wV[i] - exp_z * c) % r
                else:
                    wc = (wc - exp_z * c) % r
            exp_z = exp_z * z % r
        return wL, wR, wO, wV, wc

    def verification_msm(self, proof, g, h, G, H, rnd):
        """verifier.rs:267-449: returns (arg2 points, arg1 scalars) of the single check MSM."""
        C, r, T = self.C, self.C.r, self.transcript
        T.append_u64(b"m", len(self.V))                    # :279
        n1 = self.num_vars
        T.commit_point(b"A_I1", proof.A_I1)
        T.commit_point(b"A_O1", proof.A_O1)
        T.commit_point(b"S1", proof.S1)
        if not self.deferred:                              # create_randomized_constraints, verifier.rs:245-264
            T.r1cs_1phase_domain_sep()
        else:
            T.r1cs_2phase_domain_sep()
            callbacks, self.deferred = self.deferred, []
            self.randomizing = True
            for cb in callbacks:
                cb(self)
            self.randomizing = False
        n = self.num_vars
        n2 = n - n1
        padded_n = 1 if n == 0 else 1 << (n - 1).bit_length()
        pad = padded_n - n
        if len(G) < padded_n:                              # :297-299
            raise InvalidGeneratorsLength()
        T.commit_point(b"A_I2", proof.A_I2)
        T.commit_point(b"A_O2", proof.A_O2)
        T.commit_point(b"S2", proof.S2)
        y = T.challenge_scalar(b"y")
        z = T.challenge_scalar(b"z")
        for lab in ("T_1", "T_3", "T_4", "T_5", "T_6"):
            T.commit_point(lab.encode(), getattr(proof, lab))
        u = T.challenge_scalar(b"u")
        x = T.challenge_scalar(b"x")
        T.commit_scalar(b"t_x", proof.t_x)
        T.commit_scalar(b"t_x_blinding", proof.t_x_blinding)
        T.commit_scalar(b"e_blinding", proof.e_blinding)
        w = T.challenge_scalar(b"w")                       # :323
        wL, wR, wO, wV, wc = self.flattened_constraints(z)
        a, b = proof.ipp_proof.a, proof.ipp_proof.b
        y_inv = C.fr_inv(y)
        y_inv_vec = C.vandermonde(y_inv, padded_n)         # :342
        y_inv_wR = [wr * e % r for wr, e in zip(wR, y_inv_vec)] + [0] * pad    # :343-348
        delta = C.inner_product(y_inv_wR[:n], wL)          # :350-352
        try:
            u_sq, u_inv_sq, s = ipp_mod.verification_scalars(C, proof.ipp_proof.L, proof.ipp_proof.R, padded_n, T)
        except ipp_mod.VerificationError:
            raise R1CSError("VerificationError")
        u_for = [1] * n1 + [u] * (n2 + pad)
        g_scalars = [uo * (x * yw - a * s_i) % r for yw, uo, s_i in zip(y_inv_wR, u_for, s)]          # :368-373
        wLp, wOp = wL + [0] * pad, wO + [0] * pad
        h_scalars = [uo * (yi * (x * wl + wo - b * si) - 1) % r
                     for yi, uo, si, wl, wo in zip(y_inv_vec, u_for, reversed(s), wLp, wOp)]          # :375-390
        rr = rnd % r                                       # :392 FieldElement::random()
        x2 = x * x % r
        x3 = x * x2 % r
        r_x2 = rr * x2 % r
        rx = rr * x % r
        rx3 = rr * x3 % r
        rx4 = rx3 * x % r
        rx5 = rx4 * x % r
        rx6 = rx5 * x % r
        if getattr(self, "trace", None) is not None:       # test hook: steps 3, 5 (row a12)
            self.trace.update(y=y, z=z, u=u, x=x, w=w, n=n, n1=n1, padded_n=padded_n, wL=wL, wR=wR, wO=wO, wV=wV, wc=wc, s=s, a=a, b=b,
                              delta=delta, g_scalars=g_scalars, h_scalars=h_scalars)
        arg1 = [x, x2, x3, u * x % r, u * x2 % r, u * x3 % r]
        arg1 += [v * r_x2 % r for v in wV]                 # :416-418
        arg1 += [rx, rx3, rx4, rx5, rx6]
        arg1.append((w * (proof.t_x - a * b) + rr * (x2 * (wc + delta) - proof.t_x)) % r)            # :421
        arg1.append((-(proof.e_blinding + rr * proof.t_x_blinding)) % r)                             # :424
        arg1 += g_scalars + h_scalars + u_sq + u_inv_sq
        arg2 = [proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2]
        arg2 += self.V + [proof.T_1, proof.T_3, proof.T_4, proof.T_5, proof.T_6] + [g, h]
        arg2 += list(G[:padded_n]) + list(H[:padded_n]) + list(proof.ipp_proof.L) + list(proof.ipp_proof.R)
        return arg2, arg1

    def verify(self, proof, g, h, G, H, rnd):              # verifier.rs:267-457
        arg2, arg1 = self.verification_msm(proof, g, h, G, H, rnd)
        res = self.C.msm(arg2, arg1)                       # :451
        if not self.C.is_inf(res):
            raise R1CSError("VerificationError")


# ----------------------------------------------------------------------- gadgets
def positive_no_gadget(cs, v, n):
    """helper_constraints/positive_no.rs:8-40."""
    C = cs.C
    constraint_v = [(v.variable, C.r - 1)]
    exp_2 = 1
    for i in range(n):
        asg = None
        if v.assignment is not None:
            asg = (0, 1) if (v.assignment >> i) & 1 else (1, 0)      # :18-24 (left = 1-bit, right = bit)
        a, b, o = cs.allocate_multiplier(asg)
        cs.constrain(LC.of(o, C))                                    # :27
        cs.constrain(LC.of(a, C).add(LC.of(b, C).sub(1, C), C))      # :30  a + (b - 1)
        constraint_v.append((b, exp_2))
        exp_2 = (exp_2 + exp_2) % C.r
    cs.constrain(LC(constraint_v))                                   # :37


def bound_check_gadget(cs, v, a, b, mx, mn, n):
    """gadgets/bound_check.rs:13-39."""
    C = cs.C
    cs.constrain(LC.of(v.variable, C).sub(LC.of(mn, C), C).sub(a.variable, C))     # :26
    cs.constrain(LC.of(mx, C).sub(v.variable, C).sub(b.variable, C))               # :28
    cs.constrain(LC.of(a.variable, C).add(b.variable, C).sub(LC.of(mx - mn, C), C))  # :31 constrain_lc_with_scalar
    positive_no_gadget(cs, a, n)
    positive_no_gadget(cs, b, n)


def prove_bounded_num(prover, val, blind_v, lower, upper, bits, rng):
    """gadgets/bound_check.rs:41-92 (commitment blindings for a, b drawn from rng in order)."""
    a, b = val - lower, upper - val
    comms = []
    cv, vv = prover.commit(val, blind_v)
    comms.append(cv)
    ca, va = prover.commit(a, rng())
    comms.append(ca)
    cb, vb = prover.commit(b, rng())
    comms.append(cb)
    bound_check_gadget(prover, AllocatedQuantity(vv, val), AllocatedQuantity(va, a), AllocatedQuantity(vb, b),
                       upper, lower, bits)
    return comms


def verify_bounded_num(verifier, lower, upper, bits, commitments):
    """gadgets/bound_check.rs:94-129."""
    qs = [AllocatedQuantity(verifier.commit(c), None) for c in commitments[:3]]
    bound_check_gadget(verifier, qs[0], qs[1], qs[2], upper, lower, bits)


def shuffle_gadget(cs, xs, ys):
    """k-shuffle: {ys} is a permutation of {xs}.  The classic TWO-PHASE gadget (the example in the reference's
    constraint_system.rs:86-135 doc comments): the challenge z is drawn after the first-phase commitments, then
    prod (x_i - z) = prod (y_i - z) is enforced with 2(k-1) second-phase multipliers."""
    C = cs.C
    assert len(xs) == len(ys) and len(xs) >= 1
    k = len(xs)

    def cb(cs):
        z = cs.challenge_scalar(b"shuffle challenge")
        mz = (-z) % C.r
        if k == 1:
            cs.constrain(LC([(ys[0], 1), (xs[0], C.r - 1)]))
            return

        def chain(vs):
            _, _, out = cs.multiply(LC([(vs[k - 1], 1), (VAR_ONE, mz)]), LC([(vs[k - 2], 1), (VAR_ONE, mz)]))
            for i in range(k - 3, -1, -1):
                _, _, out = cs.multiply(LC([(out, 1)]), LC([(vs[i], 1), (VAR_ONE, mz)]))
            return out
        ox, oy = chain(xs), chain(ys)
        cs.constrain(LC([(ox, 1), (oy, C.r - 1)]))
    cs.specify_randomized_constraints(cb)


def mimc(C, xl, xr, constants, rounds):
    """gadgets/helper_constraints/mimc.rs:10-29."""
    assert len(constants) == rounds
    r = C.r
    for i in range(rounds):
        tmp1 = (xl + constants[i]) % r
        tmp2 = (tmp1 * tmp1 % r * tmp1 + xr) % r
        xr, xl = xl, tmp2
    return xl


def enforce_mimc_2_inputs(cs, left, right, rounds, constants):
    """gadgets/helper_constraints/mimc.rs:53-77 (left, right: LC); returns the LC of the image."""
    C = cs.C
    left_v, right_v = left, right
    for j in range(rounds):
        lpc = left_v.add(LC([(VAR_ONE, constants[j] % C.r)]), C)
        l, _, l_sqr = cs.multiply(lpc, LC(list(lpc.terms)))
        _, _, l_cube = cs.multiply(LC.of(l_sqr, C), LC.of(l, C))
        tmp = LC.of(l_cube, C).add(right_v, C)
        right_v, left_v = left_v, tmp
    return left_v


def poseidon_permutation(C, inp, width, frb, pr, fre, round_keys, mds, sbox):
    """gadgets/helper_constraints/poseidon.rs:202-293; sbox: "cube" | "inverse" | "quint" (:122-138); mds[j][i] as there."""
    r = C.r
    assert len(inp) == width

    def sb(e):
        if sbox == "inverse":
            return C.fr_inv(e)
        return pow(e, 3 if sbox == "cube" else 5, r)
    st = [x % r for x in inp]
    off = 0
    for rd in range(frb + pr + fre):
        full = rd < frb or rd >= frb + pr
        for i in range(width):
            st[i] = (st[i] + round_keys[off]) % r
            off += 1
            if full or i == width - 1:
                st[i] = sb(st[i])
        st = [sum(st[j] * mds[j][i] for j in range(width)) % r for i in range(width)]
    return st


def make_rng(C, seed, tag=b"blind"):
    """Deterministic blinding stream: SHAKE256(seed_le64 || tag || i_le64) mod r."""
    ctr = [0]

    def rng():
        v = C.synth_scalar(seed, ctr[0], tag)
        ctr[0] += 1
        return v
    return rng
