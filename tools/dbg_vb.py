import json, os, sys
sys.path.insert(0, "/root/repo")
import bulletproofs_amcl_b200 as bp
from oracle.curves import CURVES
from tests.util import dec_scalars
d = json.load(open("/root/repo/tests/golden/range_small.json"))["bls_m2_b8"]
ctx = bp.Context(0, 0)
C = CURVES["BLS12_381"]
m, bits, label = d["m"], d["bits"], d["label"].encode()
proof_b, comms_b = bytes.fromhex(d["proof"]), bytes.fromhex(d["commitments"])
N, lg = 16, 4
circ = bp.Circuit(ctx, bp.range_circuit_csr(0, m, bits))
dG, dH = ctx.get_generators("G", N), ctx.get_generators("H", N)
gx, hx = ctx.g1_from_msg_hash(b"g"), ctx.g1_from_msg_hash(b"h")
state = bp.r1cs_transcript_state(label)
chal = bp.r1cs_replay_challenges(0, label, proof_b, comms_b, m, lg)
key = b"k" * 13
v1, f1, var1 = circ.verify_batch(dG, dH, gx, hx, 1, proof_b, len(proof_b), comms_b, key=key, terms=True, state=state)
v2, f2, var2 = circ.verify_batch(dG, dH, gx, hx, 1, proof_b, len(proof_b), comms_b, key=key, terms=True, challenges=chal)
print(v1, v2, var1 == var2, f1 == f2)
a, b = dec_scalars(C, f1), dec_scalars(C, f2)
print([i for i in range(len(a)) if a[i] != b[i]])
y = int.from_bytes(chal[:48], "big")
print("y", hex(y))
# g[1] = x*yinv*wr1 - a*s1 ; ratio test: (g2[1]+a s1)/(g1[1]+a s1) = yinv'/yinv
from oracle import r1cs as or1cs
from oracle.merlin import Transcript
r = C.r
ch = dec_scalars(C, chal)
y, z, u, x, w = ch[:5]
uk = ch[5:]
v = or1cs.Verifier(C, Transcript(label, C))
for _ in range(m):
    or1cs.positive_no_gadget(v, or1cs.AllocatedQuantity(v.commit(C.INF), None), bits)
wL, wR, wO, wV, wc = v.flattened_constraints(z)
proof = or1cs.R1CSProof.from_bytes(C, proof_b)
aa = proof.ipp_proof.a
inv = [pow(t, -1, r) for t in uk]
s0 = 1
for t in inv:
    s0 = s0 * t % r
s1 = s0 * uk[lg - 1] * uk[lg - 1] % r
for name, arr in (("state", a), ("chal", b)):
    yi = (arr[1] + aa * s1) * pow(x * wR[1], -1, r) % r
    print(name, "implied yinv^1 =", hex(yi))
print("y^-1 =", hex(pow(y, -1, r)), "y =", hex(y))
for cand_name, cand in (("z", z), ("u", u), ("x", x), ("w", w), ("1", 1), ("0", 0)):
    print(cand_name, hex(cand), hex(pow(cand, -1, r)) if cand else "")
