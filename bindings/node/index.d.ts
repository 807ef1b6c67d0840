// Types of the zkcensus_b200 addon: the subset of snarkjs' groth16 namespace the reference uses
// (ts_inputs/src/example.ts:358-362).
export interface Groth16Proof {
  pi_a: string[];
  pi_b: string[][];
  pi_c: string[];
  protocol: "groth16";
  curve: "bn128";
}
export function fullProve(
  inputs: Record<string, unknown> | string,
  wasmPath: string,
  zkeyPath: string
): Promise<{ proof: Groth16Proof; publicSignals: string[] }>;
/** snarkjs groth16.verify(vKey, publicSignals, proof): objects or JSON strings; resolves false for an invalid proof */
export function verify(
  vKey: Record<string, unknown> | string,
  publicSignals: string[] | string,
  proof: Groth16Proof | Record<string, unknown> | string
): Promise<boolean>;
