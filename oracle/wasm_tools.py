"""Minimal MVP-wasm reader for circom-2 witness calculators (TEST INFRASTRUCTURE, not product).

Used by the oracle tooling to (1) transpile the reference's own witness calculator
(`artifacts/zkCensus/dev/160/circuit.wasm`, loaded by `zk_census_test.go:83` and
`ts_inputs/src/example.ts:360`) to C so that it can run here without a wasm runtime, and
(2) cross-check the data the product's native C++ extractor pulls out of the same file
(Poseidon constants, witness->signal map, input hashmap).

Only what circom 2.1.5 emits is handled: one memory, one table, void block types, i32/i64
integer ops, no floats/SIMD/bulk-memory.
"""
import struct


def leb_u(b, p):
    r = 0
    s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if not x & 0x80:
            return r, p


def leb_s(b, p, bits=64):
    r = 0
    s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if not x & 0x80:
            if x & 0x40:
                r -= 1 << s
            return r, p


class Func:
    __slots__ = ("idx", "type", "locals", "code_start", "code_end", "name")


class Module:
    def __init__(self, data: bytes):
        self.b = data
        assert data[:8] == b"\0asm\x01\0\0\0", "not a wasm v1 module"
        self.types = []        # (params, results)
        self.imports = []      # (module, name, typeidx)
        self.func_types = []   # typeidx of defined funcs
        self.table_elems = []  # function indices
        self.exports = {}      # name -> (kind, idx)
        self.mem_min = 0
        self.data = []         # (offset, bytes)
        self.funcs = []        # defined functions
        self.names = {}
        self._parse()

    def _parse(self):
        b = self.b
        p = 8
        n = len(b)
        while p < n:
            sid = b[p]
            p += 1
            size, p = leb_u(b, p)
            end = p + size
            if sid == 1:
                cnt, q = leb_u(b, p)
                for _ in range(cnt):
                    assert b[q] == 0x60
                    q += 1
                    np_, q = leb_u(b, q)
                    params = list(b[q:q + np_])
                    q += np_
                    nr, q = leb_u(b, q)
                    res = list(b[q:q + nr])
                    q += nr
                    self.types.append((params, res))
            elif sid == 2:
                cnt, q = leb_u(b, p)
                for _ in range(cnt):
                    l, q = leb_u(b, q)
                    mod = b[q:q + l].decode()
                    q += l
                    l, q = leb_u(b, q)
                    nm = b[q:q + l].decode()
                    q += l
                    kind = b[q]
                    q += 1
                    assert kind == 0, "only function imports expected"
                    t, q = leb_u(b, q)
                    self.imports.append((mod, nm, t))
            elif sid == 3:
                cnt, q = leb_u(b, p)
                for _ in range(cnt):
                    t, q = leb_u(b, q)
                    self.func_types.append(t)
            elif sid == 4:
                pass
            elif sid == 5:
                cnt, q = leb_u(b, p)
                flags = b[q]
                q += 1
                self.mem_min, q = leb_u(b, q)
            elif sid == 7:
                cnt, q = leb_u(b, p)
                for _ in range(cnt):
                    l, q = leb_u(b, q)
                    nm = b[q:q + l].decode()
                    q += l
                    kind = b[q]
                    q += 1
                    idx, q = leb_u(b, q)
                    self.exports[nm] = (kind, idx)
            elif sid == 9:
                cnt, q = leb_u(b, p)
                for _ in range(cnt):
                    flag, q = leb_u(b, q)
                    assert flag == 0
                    assert b[q] == 0x41
                    off, q = leb_s(b, q + 1)
                    assert b[q] == 0x0B
                    q += 1
                    ne, q = leb_u(b, q)
                    assert off == len(self.table_elems)
                    for _ in range(ne):
                        fi, q = leb_u(b, q)
                        self.table_elems.append(fi)
            elif sid == 10:
                cnt, q = leb_u(b, p)
                nimp = len(self.imports)
                for i in range(cnt):
                    bsize, q = leb_u(b, q)
                    fend = q + bsize
                    f = Func()
                    f.idx = nimp + i
                    f.type = self.func_types[i]
                    nl, q = leb_u(b, q)
                    f.locals = []
                    for _ in range(nl):
                        c, q = leb_u(b, q)
                        t = b[q]
                        q += 1
                        f.locals += [t] * c
                    f.code_start = q
                    f.code_end = fend
                    f.name = None
                    self.funcs.append(f)
                    q = fend
            elif sid == 11:
                cnt, q = leb_u(b, p)
                for _ in range(cnt):
                    flag, q = leb_u(b, q)
                    assert flag == 0
                    assert b[q] == 0x41
                    off, q = leb_s(b, q + 1)
                    assert b[q] == 0x0B
                    q += 1
                    l, q = leb_u(b, q)
                    self.data.append((off, b[q:q + l]))
                    q += l
            elif sid == 0:
                l, q = leb_u(b, p)
                nm = b[q:q + l].decode()
                q += l
                if nm == "name":
                    while q < end:
                        sub = b[q]
                        q += 1
                        ssz, q = leb_u(b, q)
                        send = q + ssz
                        if sub == 1:
                            cnt, q = leb_u(b, q)
                            for _ in range(cnt):
                                fi, q = leb_u(b, q)
                                l, q = leb_u(b, q)
                                self.names[fi] = b[q:q + l].decode()
                                q += l
                        q = send
            p = end
        nimp = len(self.imports)
        for f in self.funcs:
            f.name = self.names.get(f.idx)

    # ---- helpers -------------------------------------------------------------------------
    def func_by_name(self, name):
        for f in self.funcs:
            if f.name == name:
                return f
        raise KeyError(name)

    def memory_image(self, size=None):
        """Linear memory after data-segment initialisation (as a bytearray)."""
        top = max(o + len(d) for o, d in self.data)
        m = bytearray(size or top)
        for o, d in self.data:
            m[o:o + len(d)] = d
        return m

    def instrs(self, f):
        """Yield (pos, opcode, immediates) for function f (flat, structured ops included)."""
        b = self.b
        p = f.code_start
        end = f.code_end
        while p < end:
            pos = p
            op = b[p]
            p += 1
            imm = ()
            if op in (0x02, 0x03, 0x04):            # block loop if
                bt = b[p]
                p += 1
                imm = (bt,)
            elif op in (0x0C, 0x0D):                # br br_if
                d, p = leb_u(b, p)
                imm = (d,)
            elif op == 0x0E:                        # br_table
                cnt, p = leb_u(b, p)
                tg = []
                for _ in range(cnt + 1):
                    d, p = leb_u(b, p)
                    tg.append(d)
                imm = (tg,)
            elif op == 0x10:                        # call
                d, p = leb_u(b, p)
                imm = (d,)
            elif op == 0x11:                        # call_indirect
                t, p = leb_u(b, p)
                tb, p = leb_u(b, p)
                imm = (t, tb)
            elif 0x20 <= op <= 0x24:                # local/global get/set/tee
                d, p = leb_u(b, p)
                imm = (d,)
            elif 0x28 <= op <= 0x3E:                # loads/stores
                a, p = leb_u(b, p)
                o, p = leb_u(b, p)
                imm = (a, o)
            elif op in (0x3F, 0x40):                # memory.size/grow
                p += 1
            elif op == 0x41:
                v, p = leb_s(b, p)
                imm = (v,)
            elif op == 0x42:
                v, p = leb_s(b, p)
                imm = (v,)
            elif op in (0x43,):
                p += 4
            elif op in (0x44,):
                p += 8
            yield pos, op, imm


# ---- circom-2 specific extraction ---------------------------------------

FNV_OFF = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3


def fnv1a64(s: str) -> int:
    h = FNV_OFF
    for c in s.encode():
        h ^= c
        h = (h * FNV_PRIME) & 0xFFFFFFFFFFFFFFFF
    return h


def getter_const(mod: Module, export: str):
    """Value returned by a trivial getter `i32.const k; (return)` such as getWitnessSize."""
    kind, idx = mod.exports[export]
    f = mod.funcs[idx - len(mod.imports)]
    for pos, op, imm in mod.instrs(f):
        if op == 0x41:
            return imm[0]
    raise ValueError(export)


def fr_record(mem, addr):
    """Decode a 40-byte circom Fr record -> (shortVal, flags, long_le_int)."""
    sv, fl = struct.unpack_from("<iI", mem, addr)
    return sv, fl, int.from_bytes(mem[addr + 8:addr + 40], "little")


R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
RINV = pow(1 << 256, -1, R_MOD)


def fr_value(mem, addr):
    """Canonical value (0..r-1) of the Fr record at addr."""
    sv, fl, lv = fr_record(mem, addr)
    if fl & 0x80000000:
        if fl & 0x40000000:
            return lv * RINV % R_MOD
        return lv % R_MOD
    return sv % R_MOD


def poseidon_tables(mod: Module, mem=None):
    """{'C'|'S'|'M'|'P': {t: [Fr values]}} read from the bodies of POSEIDON_{C,S,M,P}.

    circom compiles circomlib's `POSEIDON_C(t)` etc. (poseidon_constants.circom, included by
    `circuit/census.circom:3`) into `if Fr_eq(t, const) { Fr_copy(dst + 40*k, constAddr) ... }`
    chains; element k-1 of the returned array is the constant at `constAddr`.
    """
    mem = mem or mod.memory_image()
    fr_eq = next(i for i, n in mod.names.items() if n == "Fr_eq")
    fr_copy = next(i for i, n in mod.names.items() if n == "Fr_copy")
    out = {}
    for key in "CSMP":
        f = next(f for f in mod.funcs if f.name and f.name.startswith(f"POSEIDON_{key}_"))
        ins = list(mod.instrs(f))
        tabs = {}
        cur = None
        for i, (pos, op, imm) in enumerate(ins):
            if op != 0x10:
                continue
            if imm[0] == fr_eq:
                assert ins[i - 1][1] == 0x41
                cur = fr_value(mem, ins[i - 1][2][0])
                tabs[cur] = {}
            elif imm[0] == fr_copy and cur is not None:
                # i32.const k; i32.const 40; i32.mul; local.get; i32.add; i32.const addr; call
                if (ins[i - 1][1] == 0x41 and ins[i - 2][1] == 0x6A and ins[i - 3][1] == 0x20 and
                        ins[i - 4][1] == 0x6C and ins[i - 5][1] == 0x41 and ins[i - 5][2][0] == 40 and
                        ins[i - 6][1] == 0x41):
                    k = ins[i - 6][2][0]
                    tabs[cur][k - 1] = fr_value(mem, ins[i - 1][2][0])
        res = {}
        for t, d in tabs.items():
            if d:
                n = max(d) + 1
                assert sorted(d) == list(range(n)), (key, t)
                res[t] = [d[i] for i in range(n)]
        out[key] = res
    return out


def witness_map(mod: Module, mem=None):
    """witness index -> circom signal index (u32 table)."""
    mem = mem or mod.memory_image()
    nw = getter_const(mod, "getWitnessSize")
    # the table is the data segment that starts right after the input-set flags; locate it as the
    # segment whose length is exactly 4*nw
    for off, d in mod.data:
        if len(d) == 4 * nw:
            return list(struct.unpack_from(f"<{nw}I", d, 0)), off
    raise ValueError("witness map segment not found")


def input_hashmap(mod: Module, names, mem=None):
    """{name: (signalPos, size)} from the 256-slot hashmap (mem 2032)."""
    mem = mem or mod.memory_image()
    # locate: the segment holding 256 x 16-byte slots
    base = None
    for off, d in mod.data:
        if len(d) == 4096:
            base = off
    assert base is not None
    out = {}
    for nm in names:
        h = fnv1a64(nm)
        slot = h & 0xFF
        for _ in range(256):
            hh, pos, size = struct.unpack_from("<QII", mem, base + 16 * slot)
            if hh == h:
                out[nm] = (pos, size)
                break
            slot = (slot + 1) & 0xFF
        else:
            raise KeyError(nm)
    return out
