"""B200-native Groth16 prover for Vocdoni's census.circom (drop-in for the proving path of
vocdoni/zk-franchise-proof-circuit).  See DESIGN.md."""
