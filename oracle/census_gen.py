"""Synthetic census / SIK trees and circuit inputs (TEST + BENCH INFRASTRUCTURE, not product).

Restates the reference's input generation: `internal/inputs.go:33-98` (MockInputs), `internal/helpers.go:
17-85` (BigToFF, BytesToArbo, GenTree with vocdoni's arbo Poseidon tree, un-vendored: `go.mod:7`) and its
TypeScript twin `ts_inputs/src/inputs.ts:55-89`, following SURVEY.md section 8d "Config 2":

  address_i  = first 20 bytes of sha256(seed || "addr" || i)      (key = little-endian int, arbo.BytesToBigInt)
  signature_i= 64 bytes sha256-CTR(seed || "sig" || i) as big-endian int mod r   (inputs.go:92)
  password   = "password123" big-endian mod r (inputs.go:41,91); availableWeight 10, voteWeight 5 (:34,94)
  electionId = the reference's hex (inputs.go:60) -> sha256 -> two LE 128-bit halves (helpers.go:28-34)
  census tree: arbo sparse Merkle tree, Poseidon, LSB-first key bits: empty -> 0, single leaf ->
     H(key, value, 1), else H(left, right); siblings of a key = other branch at each depth until alone
  SIK tree: key = address, value = Poseidon3(address, password, signature)   (census.circom:74-77)
  siblings padded with zeros to nLevels + 1 (inputs.go:52,72)

Pinned by: Poseidon constants reproduce the fixture's nullifier / roots (tests/test_oracle_witness.py),
and every generated input passes all asserts of the reference wasm.
"""
import hashlib

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]
ELECTION_HEX = "7faeab7a7d250527d614e952ae8e446825bd1124c6def410844c7c383d1519a6"


class Poseidon:
    """circomlib optimised Poseidon (same schedule as poseidon.circom PoseidonEx)."""

    def __init__(self, tables):
        self.t = tables

    def __call__(self, inputs):
        t = len(inputs) + 1
        C, S, M, Pm = (self.t[k][t] for k in "CSMP")
        RP = N_ROUNDS_P[t - 2]
        st = [0] + [x % P for x in inputs]
        st = [(st[j] + C[j]) % P for j in range(t)]
        mix = lambda st, Mx: [sum(Mx[j * t + i] * st[j] for j in range(t)) % P for i in range(t)]
        for r in range(3):
            st = [pow(x, 5, P) for x in st]
            st = [(st[j] + C[(r + 1) * t + j]) % P for j in range(t)]
            st = mix(st, M)
        st = [pow(x, 5, P) for x in st]
        st = [(st[j] + C[4 * t + j]) % P for j in range(t)]
        st = mix(st, Pm)
        for r in range(RP):
            st[0] = (pow(st[0], 5, P) + C[5 * t + r]) % P
            base = (2 * t - 1) * r
            o0 = sum(S[base + i] * st[i] for i in range(t)) % P
            for i in range(1, t):
                st[i] = (st[i] + st[0] * S[base + t + i - 1]) % P
            st[0] = o0
        for r in range(3):
            st = [pow(x, 5, P) for x in st]
            st = [(st[j] + C[5 * t + RP + r * t + j]) % P for j in range(t)]
            st = mix(st, M)
        st = [pow(x, 5, P) for x in st]
        return sum(M[j * t] * st[j] for j in range(t)) % P


def bytes_to_arbo(b: bytes):
    h = hashlib.sha256(b).digest()
    return [int.from_bytes(h[:16], "little"), int.from_bytes(h[16:], "little")]


class SMT:
    """arbo-style sparse Merkle tree over (key, value) pairs; keys are ints, path = LSB-first bits."""

    def __init__(self, H, leaves: dict):
        self.H = H
        self.leaves = leaves
        self.memo = {}
        self.root = self._node(tuple(sorted(leaves)), 0)

    def _node(self, keys, depth):
        if not keys:
            return 0
        if len(keys) == 1:
            k = keys[0]
            return self.H([k, self.leaves[k], 1])
        memo_key = (keys[0], len(keys), depth)
        if memo_key in self.memo:
            return self.memo[memo_key]
        left = tuple(k for k in keys if not (k >> depth) & 1)
        right = tuple(k for k in keys if (k >> depth) & 1)
        # an empty side hashes as 0; arbo keeps descending until the keys split
        h = self.H([self._node(left, depth + 1), self._node(right, depth + 1)])
        self.memo[memo_key] = h
        return h

    def siblings(self, key):
        keys = tuple(sorted(self.leaves))
        out = []
        depth = 0
        while len(keys) > 1:
            left = tuple(k for k in keys if not (k >> depth) & 1)
            right = tuple(k for k in keys if (k >> depth) & 1)
            if (key >> depth) & 1:
                out.append(self._node(left, depth + 1))
                keys = right
            else:
                out.append(self._node(right, depth + 1))
                keys = left
            depth += 1
        return out


def gen_census(tables, n_voters, seed=0xC0FFEE, n_levels=160, available_weight=10, vote_weight=5):
    """Returns a list of n_voters inputs dicts (decimal strings, same 12 keys as inputs_example.json)."""
    H = Poseidon(tables)
    sd = seed.to_bytes(8, "big")
    password = int.from_bytes(b"password123", "big") % P
    election = bytes_to_arbo(bytes.fromhex(ELECTION_HEX))
    vote_hash = bytes_to_arbo(available_weight.to_bytes(1, "big"))
    voters = []
    for i in range(n_voters):
        ib = i.to_bytes(4, "big")
        addr = int.from_bytes(hashlib.sha256(sd + b"addr" + ib).digest()[:20], "little")
        sig = b"".join(hashlib.sha256(sd + b"sig" + ib + bytes([c])).digest() for c in range(2))
        sig = int.from_bytes(sig, "big") % P
        sik = H([addr, password, sig])
        voters.append((addr, sig, sik))
    census = SMT(H, {a: available_weight for a, _, _ in voters})
    siktree = SMT(H, {a: k for a, _, k in voters})
    out = []
    for addr, sig, sik in voters:
        cs = census.siblings(addr)
        ss = siktree.siblings(addr)
        pad = lambda s: [str(x) for x in s] + ["0"] * (n_levels + 1 - len(s))
        out.append({
            "electionId": [str(election[0]), str(election[1])],
            "nullifier": str(H([sig, password, election[0], election[1]])),
            "availableWeight": str(available_weight),
            "voteHash": [str(vote_hash[0]), str(vote_hash[1])],
            "sikRoot": str(siktree.root),
            "censusRoot": str(census.root),
            "address": str(addr),
            "password": str(password),
            "signature": str(sig),
            "voteWeight": str(vote_weight),
            "censusSiblings": pad(cs),
            "sikSiblings": pad(ss),
        })
    return out
