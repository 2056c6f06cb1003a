"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under this package is product code.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import, link or execute it, and there only as
the checker (or as the timed CPU baseline), never as the thing that is shipped.  The product path
(`zkp_b200/`) never imports this package and fails loudly when its CUDA library is missing.

What is restated, and from where
--------------------------------
The reference (`dalek-cryptography/zkp` @ 419c9bf) is pure Rust and delegates all arithmetic on the hot
path to third-party crates whose source is NOT under /root/reference:

  * curve25519-dalek  semver "^2" (Cargo.toml:27)  => 2.1.3 is the newest matching release
  * merlin            semver "^2" (Cargo.toml:21)  => 2.0.1

so the algorithms are restated from their published form (RFC 9496 for ristretto255, the dalek 2.x
`scalar.rs`/`straus.rs`/`pippenger.rs` algorithms, merlin 2.0 `strobe.rs`/`transcript.rs`) and anchored on the
reference's own call sites:

  MSM         src/toolbox/prover.rs:94, src/toolbox/verifier.rs:97, :162, src/toolbox/batch_verifier.rs:219
  compress    src/toolbox/mod.rs:180, :204
  decompress  src/toolbox/verifier.rs:90, :164, src/toolbox/batch_verifier.rs:226
  identity    src/toolbox/mod.rs:191, :215, src/toolbox/verifier.rs:168, src/toolbox/batch_verifier.rs:230
  transcript  src/toolbox/mod.rs:165-228

Pinning status
--------------
The reference holds NO golden vectors / known-answer tests for this path (all of tests/*.rs assert only
is_ok()/is_err() on self-generated randomised proofs) and cannot be built here (no rustc/cargo, no crate
sources).  With respect to *reference-produced bytes* parity is therefore UNPINNED, and DESIGN.md says so.
What pins the oracle instead (tests/test_oracle_*.py, all run on CPU):

  * RFC 9496 Appendix A vectors (multiples of the generator, invalid encodings, hash-to-group) -- the same
    vectors curve25519-dalek's own ristretto.rs unit tests carry;
  * libsodium 1.0.20's independent ristretto255 implementation (found inside the pyzmq wheel), compared on
    random scalar mults, additions, decode validity of random strings, hash-to-group;
  * the Merlin cross-implementation conformance vector;
  * ristretto255 encodings and scalars mod l are canonical, so any correct implementation is byte-identical
    to the reference on MSM outputs, challenges and accept bits.

Modules: field25519 / ristretto (group), scalar (mod l + dalek recodings), msm (dalek's Straus CT, Straus
vartime NAF-5, Pippenger), merlin (Keccak-f, STROBE-128, Transcript, TranscriptRng), toolbox (Prover /
Verifier / BatchVerifier flows), sodium (libsodium loader for cross-checks), c/ (C restatement of the
serial u64 backend and, in ref_ifma.h, of the 4-way vector `simd_backend` = the timed CPU baselines).
"""
