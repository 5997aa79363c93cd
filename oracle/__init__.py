"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain-Python (big-int) and plain-C restatement of the arithmetic that
lovesh/bulletproofs-amcl delegates to `amcl_wrapper` / `miracl_amcl` / `merlin`
for its data-parallel hot path (G1 MSM, IPP, R1CS prover/verifier MSMs and Fr
vector algebra).

PARITY STATUS: **parity unpinned** by the reference itself.  The reference
(`/root/reference`) holds no golden vectors (every test is a prove->verify round
trip with fresh randomness), its arithmetic dependency
`amcl_wrapper ^0.1.5` is not vendored, there is no Cargo.lock and no Rust
toolchain here.  What this oracle IS pinned against (tests/test_oracle_*.py):
  * SHAKE256 / Keccak-f[1600] against hashlib;
  * Merlin v1.0's published `equivalence_simple` known answer;
  * the standard BLS12-381 G1 generator, curve order and cofactor identities;
  * algebraic invariants (completeness of prove->verify, padding invariance
    ipp.rs:431-471, MSM linearity) and Python-vs-C cross checks.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this package.  The product
(`bulletproofs-amcl_b200/`) never does.
"""
