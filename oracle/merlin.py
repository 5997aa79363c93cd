"""Oracle: Keccak-f[1600], STROBE-128 and Merlin v1.0 transcripts (test infrastructure).

The reference takes its Fiat-Shamir transcript from the un-vendored crate
`merlin = "1"` (Cargo.toml:10) and drives it through `TranscriptProtocol`
(`/root/reference/src/transcript.rs:29-61`).  This file restates the published
STROBE-128/1.0.2 + Merlin v1.0 framing and is pinned by
Merlin's own `equivalence_simple` known answer (tests/test_oracle_merlin.py).
"""

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M if n else x


def keccak_f1600(state: bytearray):
    """In-place Keccak-f[1600] on a 200-byte state (lanes little endian)."""
    A = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = (A[x][y] & _M).to_bytes(8, "little")


def shake256(msg: bytes, outlen: int) -> bytes:
    """SHAKE256 built on keccak_f1600 (only used to validate the permutation against hashlib)."""
    rate = 136
    st = bytearray(200)
    m = bytearray(msg) + b"\x1f"
    while len(m) % rate:
        m.append(0)
    m[-1] |= 0x80
    for off in range(0, len(m), rate):
        for i in range(rate):
            st[i] ^= m[off + i]
        keccak_f1600(st)
    out = b""
    while len(out) < outlen:
        out += bytes(st[:rate])
        if len(out) < outlen:
            keccak_f1600(st)
    return out[:outlen]


class Strobe128:
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.st, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[self.R + 1] ^= 0x80
        keccak_f1600(self.st)
        self.pos = self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.st[self.pos] ^= b
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.st[self.pos])
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert not flags & self.FLAG_T
        old = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin::Transcript (v1.0 framing) + the reference's TranscriptProtocol extension."""

    def __init__(self, label: bytes, curve=None):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.curve = curve
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, msg: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(msg).to_bytes(4, "little"), True)
        self.strobe.ad(msg, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, x.to_bytes(8, "little"))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n)

    # ---- TranscriptProtocol (transcript.rs:29-61)
    def innerproduct_domain_sep(self, n):          # transcript.rs:30-33
        self.append_message(b"dom-sep", b"ipp v1")
        self.append_u64(b"n", n)

    def r1cs_domain_sep(self):                     # transcript.rs:35-37
        self.append_message(b"dom-sep", b"r1cs v1")

    def r1cs_1phase_domain_sep(self):              # transcript.rs:39-41
        self.append_message(b"dom-sep", b"r1cs-1phase")

    def r1cs_2phase_domain_sep(self):              # transcript.rs:43-45
        self.append_message(b"dom-sep", b"r1cs-2phase")

    def commit_scalar(self, label, s):             # transcript.rs:47-49
        self.append_message(label, self.curve.fr_to_bytes(s))

    def commit_point(self, label, P):              # transcript.rs:51-53
        self.append_message(label, self.curve.g1_to_bytes(P))

    def challenge_scalar(self, label):             # transcript.rs:55-60
        return self.curve.fr_from_bytes(self.challenge_bytes(label, self.curve.MODBYTES))
