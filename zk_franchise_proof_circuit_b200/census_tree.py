"""Synthetic census / SIK trees and circuit inputs with the hashing done on the GPU.

The step before the proving path in the reference: `internal/inputs.go:33-98` (MockInputs) builds two arbo
Poseidon sparse Merkle trees with `internal/helpers.go:36-85` (GenTree) and marshals the 12-key inputs JSON;
`ts_inputs/src/inputs.ts:55-89` is its TypeScript twin.  Here the trees of a whole census are built level by
level with one batched Poseidon launch per level (zkb_poseidon_hash), so a batch of thousands of voters is not
bound by host hashing.  Follows SURVEY.md 8d "Config 2":

  address_i   = first 20 bytes of sha256(seed || "addr" || i), key = little-endian int (arbo.BytesToBigInt)
  signature_i = 64 bytes sha256-CTR(seed || "sig" || i), big-endian int mod r      (inputs.go:92)
  password    = "password123" big-endian mod r (inputs.go:41,91); availableWeight 10, voteWeight 5 (:34,94)
  electionId  = the reference's hex (inputs.go:60) -> sha256 -> two little-endian 128-bit halves (helpers.go:28-34)
  tree        = arbo semantics: path = key bits LSB first; empty -> 0; single key -> H(key, value, 1);
                otherwise H(left, right); siblings of a key = the other branch at every depth until it is alone
  SIK value   = Poseidon3(address, password, signature) (census.circom:74-77); siblings zero-padded to nLevels+1
"""
import hashlib

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ELECTION_HEX = "7faeab7a7d250527d614e952ae8e446825bd1124c6def410844c7c383d1519a6"


def bytes_to_arbo(b: bytes):
    """internal/helpers.go:28-34 BytesToArbo / ts_inputs arbo_utils.ts:22-33 toHash"""
    h = hashlib.sha256(b).digest()
    return [int.from_bytes(h[:16], "little"), int.from_bytes(h[16:], "little")]


class _Node:
    __slots__ = ("keys", "depth", "left", "right", "hash")


class SparseMerkleTree:
    """arbo-style tree over {key: value}; all hashing goes through `hasher(rows) -> ints` in per-depth batches.
    Python statement of the tree semantics, kept as the cross-check of the library's builder (Circuit.census_tree,
    zkb_census_tree), which gen_census uses."""

    def __init__(self, hasher, leaves: dict):
        self.leaves = leaves
        root = self._build(sorted(leaves), 0)
        by_depth = {}
        stack = [root] if root else []
        leaf_nodes = []
        while stack:
            nd = stack.pop()
            if len(nd.keys) == 1:
                leaf_nodes.append(nd)
                continue
            by_depth.setdefault(nd.depth, []).append(nd)
            for ch in (nd.left, nd.right):
                if ch is not None:
                    stack.append(ch)
        hs = hasher([(nd.keys[0], leaves[nd.keys[0]], 1) for nd in leaf_nodes])
        for nd, h in zip(leaf_nodes, hs):
            nd.hash = h
        for depth in sorted(by_depth, reverse=True):
            nodes = by_depth[depth]
            hs = hasher([(nd.left.hash if nd.left else 0, nd.right.hash if nd.right else 0) for nd in nodes])
            for nd, h in zip(nodes, hs):
                nd.hash = h
        self.root_node = root
        self.root = root.hash if root else 0

    def _build(self, keys, depth):
        if not keys:
            return None
        nd = _Node()
        nd.keys, nd.depth, nd.left, nd.right, nd.hash = keys, depth, None, None, 0
        if len(keys) > 1:
            nd.left = self._build([k for k in keys if not (k >> depth) & 1], depth + 1)
            nd.right = self._build([k for k in keys if (k >> depth) & 1], depth + 1)
        return nd

    def siblings(self, key):
        out = []
        nd = self.root_node
        while nd is not None and len(nd.keys) > 1:
            if (key >> nd.depth) & 1:
                out.append(nd.left.hash if nd.left else 0)
                nd = nd.right
            else:
                out.append(nd.right.hash if nd.right else 0)
                nd = nd.left
        return out


def gen_census(circuit, n_voters, seed=0xC0FFEE, n_levels=160, available_weight=10, vote_weight=5):
    """n_voters inputs dicts (decimal strings, the 12 keys of inputs_example.json) for one synthetic census.
    `circuit` is a loaded prover.Circuit (its Poseidon constants are the circuit's own)."""
    H = circuit.poseidon
    sd = seed.to_bytes(8, "big")
    password = int.from_bytes(b"password123", "big") % R_MOD
    election = bytes_to_arbo(bytes.fromhex(ELECTION_HEX))
    vote_hash = bytes_to_arbo(available_weight.to_bytes(1, "big"))
    addrs, sigs = [], []
    for i in range(n_voters):
        ib = i.to_bytes(4, "big")
        addrs.append(int.from_bytes(hashlib.sha256(sd + b"addr" + ib).digest()[:20], "little"))
        sig = b"".join(hashlib.sha256(sd + b"sig" + ib + bytes([c])).digest() for c in range(2))
        sigs.append(int.from_bytes(sig, "big") % R_MOD)
    siks = H([(a, password, s) for a, s in zip(addrs, sigs)])
    nulls = H([(s, password, election[0], election[1]) for s in sigs])
    # both trees by the library's builder (zkb_census_tree: trie layout in C++, hashing on the GPU)
    census_root, census_sibs = circuit.census_tree({a: available_weight for a in addrs}, n_levels)
    sik_root, sik_sibs = circuit.census_tree(dict(zip(addrs, siks)), n_levels)

    class _T:
        def __init__(self, root, sibs):
            self.root, self._s = root, sibs

        def siblings(self, key):
            return self._s[key]
    census, siktree = _T(census_root, census_sibs), _T(sik_root, sik_sibs)
    out = []
    pad = lambda s: [str(x) for x in s] + ["0"] * (n_levels + 1 - len(s))
    for a, s, nul in zip(addrs, sigs, nulls):
        out.append({
            "electionId": [str(election[0]), str(election[1])],
            "nullifier": str(nul),
            "availableWeight": str(available_weight),
            "voteHash": [str(vote_hash[0]), str(vote_hash[1])],
            "sikRoot": str(siktree.root),
            "censusRoot": str(census.root),
            "address": str(a),
            "password": str(password),
            "signature": str(s),
            "voteWeight": str(vote_weight),
            "censusSiblings": pad(census.siblings(a)),
            "sikSiblings": pad(siktree.siblings(a)),
        })
    return out


def tree_depth(inputs: dict) -> int:
    """Depth of a voter's Merkle path as the circuit sees it: index of the last non-zero sibling + 1, maximum over the
    census and SIK trees (SMTLevIns; the levels below are the proof-independent Poseidon2(0,0) blocks of SURVEY 8a W7)."""
    d = 0
    for k in ("censusSiblings", "sikSiblings"):
        nz = [i for i, x in enumerate(inputs[k]) if int(x) != 0]
        d = max(d, nz[-1] + 1 if nz else 0)
    return d


def gen_census_depth(circuit, n_voters, depth, seed=0xC0FFEE, n_levels=160, available_weight=10, vote_weight=5):
    """n_voters inputs dicts whose Merkle paths have exactly `depth` non-zero siblings in both trees (1 <= depth <=
    n_levels): every voter sits in its own comb-shaped census of depth + 1 leaves - the voter's address plus, for every
    level j < depth, the address with bit j flipped, which splits off alone at level j (arbo semantics as in
    SparseMerkleTree).  The sibling at level j is then that leaf's hash and the voter's own leaf sits at level `depth`.
    All voters' trees are hashed together, one batched Poseidon launch per level (zkb_poseidon_hash).
    Used by bench.py to measure throughput as a function of the tree depth (SURVEY.md 8a W7 measurement rule):
    depth 4 is the reference's own 10-leaf shape (internal/helpers.go:56-61), depth 160 the circuit's maximum."""
    assert 1 <= depth <= n_levels
    H = circuit.poseidon
    sd = seed.to_bytes(8, "big")
    password = int.from_bytes(b"password123", "big") % R_MOD
    election = bytes_to_arbo(bytes.fromhex(ELECTION_HEX))
    vote_hash = bytes_to_arbo(available_weight.to_bytes(1, "big"))
    addrs, sigs = [], []
    for i in range(n_voters):
        ib = i.to_bytes(4, "big") + depth.to_bytes(2, "big")
        addrs.append(int.from_bytes(hashlib.sha256(sd + b"deep-addr" + ib).digest()[:20], "little"))
        sig = b"".join(hashlib.sha256(sd + b"deep-sig" + ib + bytes([c])).digest() for c in range(2))
        sigs.append(int.from_bytes(sig, "big") % R_MOD)
    nulls = H([(s, password, election[0], election[1]) for s in sigs])
    sik_own = H([(a, password, s) for a, s in zip(addrs, sigs)])
    # the flipped-bit neighbours: census value = weight, SIK value = an arbitrary non-zero field element
    nb_keys = [[a ^ (1 << j) for j in range(depth)] for a in addrs]
    nb_sik = [[int.from_bytes(hashlib.sha256(sd + b"nb-sik" + k.to_bytes(20, "little")).digest(), "big") % R_MOD
               for k in ks] for ks in nb_keys]
    out = []
    roots, sibs = {}, {}
    for name, own_val, nb_val in (("census", [available_weight] * n_voters, None), ("sik", sik_own, nb_sik)):
        leaf_rows = []
        for v in range(n_voters):
            leaf_rows.append((addrs[v], own_val[v], 1))
            for j in range(depth):
                leaf_rows.append((nb_keys[v][j], available_weight if nb_val is None else nb_val[v][j], 1))
        lh = H(leaf_rows)
        own = [lh[v * (depth + 1)] for v in range(n_voters)]
        sib = [[lh[v * (depth + 1) + 1 + j] for j in range(depth)] for v in range(n_voters)]
        node = own
        for j in range(depth - 1, -1, -1):        # node at level j = H(left, right), the voter's side chosen by bit j
            rows = [((sib[v][j], node[v]) if (addrs[v] >> j) & 1 else (node[v], sib[v][j])) for v in range(n_voters)]
            node = H(rows)
        roots[name], sibs[name] = node, sib
    pad = lambda s: [str(x) for x in s] + ["0"] * (n_levels + 1 - len(s))
    for v in range(n_voters):
        out.append({
            "electionId": [str(election[0]), str(election[1])],
            "nullifier": str(nulls[v]),
            "availableWeight": str(available_weight),
            "voteHash": [str(vote_hash[0]), str(vote_hash[1])],
            "sikRoot": str(roots["sik"][v]),
            "censusRoot": str(roots["census"][v]),
            "address": str(addrs[v]),
            "password": str(password),
            "signature": str(sigs[v]),
            "voteWeight": str(vote_weight),
            "censusSiblings": pad(sibs["census"][v]),
            "sikSiblings": pad(sibs["sik"][v]),
        })
    return out
