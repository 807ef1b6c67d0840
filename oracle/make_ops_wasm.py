#!/usr/bin/env python3
"""A THIRD circom-runtime program for the generic witness path (TEST INFRASTRUCTURE): every field operation of the
circom runtime, a nested conditional, a one-armed conditional and an assert, on three inputs.

The census program uses a third of the runtime's operations and no nested conditionals; no circom compiler exists in
the image.  This assembles the body of a small template by hand in the code shapes circom 2.1.5 itself emits (taken
from the reference wasm: `AND_156_run` for an operator, `IsZero_79_run` for `if / else` and for a `===` check) and
splices it into the reference wasm in place of `AND_156_run`; `AND` becomes the main component.  The field runtime
(`Fr_*`), `init`, `setInputSignal`, `getWitness` ... are the reference's own code.

    template OpsTest() {
        signal input x[3];  signal output o[26];
        o[0] <-- x[0] \\ x[1];   o[1] <-- x[0] % x[1];    o[2] <-- x[0] ** x[2];   o[3] <-- x[0] / x[1];
        o[4] <-- x[0] | x[1];    o[5] <-- x[0] ^ x[1];    o[6] <-- ~x[0];          o[7] <-- x[0] << x[2];
        o[8] <-- x[0] >> x[2];   o[9] <-- x[0] < x[1];    o[10] <-- x[0] > x[1];   o[11] <-- x[0] <= x[1];
        o[12] <-- x[0] >= x[1];  o[13] <-- x[0] == x[1];  o[14] <-- x[0] != x[1];  o[15] <-- x[0] && x[2];
        o[16] <-- x[0] || x[2];  o[17] <-- !x[2];         o[18] <-- -x[0];         o[19] <-- x[0] & x[1];
        o[20] <-- 1 / x[1] (Fr_inv);   o[21] <-- x[0] * x[1] + x[2];   o[22] <-- x[0] - x[1];
        if (x[0] < x[1]) { if (x[2] != 0) { o[23] <-- x[0] * x[2]; } else { o[23] <-- x[1]; } }
        else { o[23] <-- x[0] - x[1]; }
        var m = x[0];  if (x[0] > x[1]) { m = x[1]; }  o[24] <-- m;           // one-armed: min in the signed order
        o[25] <-- x[2] * x[2];
        assert(x[2] <= x[1]);
    }

Signals of the main component (circom order: outputs, then inputs): o[k] = signal 1 + k, x[j] = signal 27 + j; the
witness is the first 30 signals (identity map).

    python oracle/make_ops_wasm.py out.wasm
"""
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import wasm_tools as W            # noqa: E402
import make_smt_wasm as SW        # noqa: E402

N_OUT, N_IN = 26, 3
N_WIRES = 1 + N_OUT + N_IN
INPUTS = [("x", N_IN)]


def leb_u(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def leb_s(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        done = (v == 0 and not (b & 0x40)) or (v == -1 and (b & 0x40))
        out.append(b | (0 if done else 0x80))
        if done:
            return bytes(out)


class Asm:
    """the handful of instructions circom's template code consists of"""

    def __init__(self, fn):
        self.b = bytearray()
        self.fn = fn                     # name -> function index

    def const(self, v): self.b += b"\x41" + leb_s(v)
    def get(self, i): self.b += b"\x20" + leb_u(i)
    def set(self, i): self.b += b"\x21" + leb_u(i)
    def add(self): self.b += b"\x6a"
    def mul(self): self.b += b"\x6c"
    def eqz(self): self.b += b"\x45"
    def load(self): self.b += b"\x28\x02\x00"
    def store(self): self.b += b"\x36\x02\x00"
    def call(self, name): self.b += b"\x10" + leb_u(self.fn[name] if isinstance(name, str) else name)
    def if_(self): self.b += b"\x04\x40"
    def else_(self): self.b += b"\x05"
    def end(self): self.b += b"\x0b"
    def ret(self): self.b += b"\x0f"

    # operand addresses, as circom writes them
    def sig(self, k):                    # &signal[k] of this component: k * 40 + signal_start (local 2)
        self.const(k); self.const(40); self.mul(); self.get(2); self.add()

    def tmp(self, j):                    # expression-stack slot j: local 7 + 40 j
        self.get(7); self.const(40 * j); self.add()

    def operand(self, o):
        kind, v = o
        {"sig": self.sig, "tmp": self.tmp, "const": self.const}[kind](v)

    def op(self, name, dst, *src):       # dst = name(src...) through a temporary, then Fr_copy (circom's shape)
        self.operand(dst)
        self.tmp(0)
        for s in src:
            self.operand(s)
        self.call(name)
        self.tmp(0)
        self.call("Fr_copy")

    def op_into(self, name, dst, *src):  # dst = name(src...) directly (sub-expressions on the expression stack)
        self.operand(dst)
        for s in src:
            self.operand(s)
        self.call(name)

    def cond(self, name, a, b):          # leaves the i32 truth value of `a name b` on the wasm stack
        self.op_into(name, ("tmp", 0), a, b)
        self.tmp(0)
        self.call("Fr_isTrue")


def ops_body(fn, zero_addr, one_addr):
    a = Asm(fn)
    O = lambda k: ("sig", k)                      # outputs first
    X = lambda j: ("sig", N_OUT + j)              # then inputs
    # prologue (AND_156_run / IsZero_79_run)
    a.const(0); a.load(); a.set(6)
    a.get(6); a.const(0); a.add(); a.set(7)
    a.const(40 * 4); a.call("reserveStackFr"); a.set(1)
    a.get(0); a.const(4); a.add(); a.load(); a.set(2)
    binary = [("Fr_idiv", 0, 0, 1), ("Fr_mod", 1, 0, 1), ("Fr_pow", 2, 0, 2), ("Fr_div", 3, 0, 1), ("Fr_bor", 4, 0, 1),
              ("Fr_bxor", 5, 0, 1), ("Fr_shl", 7, 0, 2), ("Fr_shr", 8, 0, 2), ("Fr_lt", 9, 0, 1), ("Fr_gt", 10, 0, 1),
              ("Fr_leq", 11, 0, 1), ("Fr_geq", 12, 0, 1), ("Fr_eq", 13, 0, 1), ("Fr_neq", 14, 0, 1),
              ("Fr_land", 15, 0, 2), ("Fr_lor", 16, 0, 2), ("Fr_band", 19, 0, 1), ("Fr_sub", 22, 0, 1),
              ("Fr_mul", 25, 2, 2)]
    for name, k, i, j in binary:
        a.op(name, O(k), X(i), X(j))
    for name, k, i in [("Fr_bnot", 6, 0), ("Fr_lnot", 17, 2), ("Fr_neg", 18, 0), ("Fr_inv", 20, 1)]:
        a.op(name, O(k), X(i))
    # o[21] = x0 * x1 + x2: the product on expression slot 1, the sum on slot 0
    a.operand(O(21)); a.tmp(0)
    a.op_into("Fr_mul", ("tmp", 1), X(0), X(1))
    a.tmp(1); a.operand(X(2)); a.call("Fr_add")
    a.tmp(0); a.call("Fr_copy")
    # nested two-armed conditionals
    a.cond("Fr_lt", X(0), X(1))
    a.if_()
    a.cond("Fr_neq", X(2), ("const", zero_addr))
    a.if_()
    a.op("Fr_mul", O(23), X(0), X(2))
    a.else_()
    a.operand(O(23)); a.operand(X(1)); a.call("Fr_copy")
    a.end()
    a.else_()
    a.op("Fr_sub", O(23), X(0), X(1))
    a.end()
    # one-armed conditional on a variable (expression slot 3 plays the var)
    a.tmp(3); a.operand(X(0)); a.call("Fr_copy")
    a.cond("Fr_gt", X(0), X(1))
    a.if_()
    a.tmp(3); a.operand(X(1)); a.call("Fr_copy")
    a.end()
    a.operand(O(24)); a.tmp(3); a.call("Fr_copy")
    # assert(x2 <= x1): IsZero_79_run's `===` epilogue
    a.cond("Fr_leq", X(2), X(1))
    a.eqz()
    a.if_()
    a.const(79); a.const(33); a.call("buildBufferMessage"); a.call(1)      # import 1 = printErrorMessage
    a.const(4); a.ret()
    a.end()
    # epilogue
    a.const(0); a.get(1); a.store()
    a.const(0); a.end()
    return bytes(a.b)


def toint_body(fn):
    """a template that uses a signal as an integer (`var n = x[0]; for (i < n) ...`): circom calls Fr_toInt on it.  The
    generic extractor cannot turn that into a straight-line program and must say so."""
    a = Asm(fn)
    a.const(0); a.load(); a.set(6)
    a.get(6); a.const(0); a.add(); a.set(7)
    a.const(40 * 4); a.call("reserveStackFr"); a.set(1)
    a.get(0); a.const(4); a.add(); a.load(); a.set(2)
    a.sig(N_OUT); a.call("Fr_toInt"); a.set(3)            # local 3 = x[0] as an integer
    a.const(0); a.get(1); a.store()
    a.const(0); a.end()
    return bytes(a.b)


def replace_body(wasm: bytes, code_index: int, instrs: bytes) -> bytes:
    """new module in which the code_index-th body of the code section keeps its locals and gets `instrs`"""
    p = 8
    while p < len(wasm):
        sid = wasm[p]
        size, q = W.leb_u(wasm, p + 1)
        if sid == 10:
            sec_start, sec_end = q, q + size
            cnt, r = W.leb_u(wasm, q)
            bodies_at = r
            for i in range(cnt):
                bsize, r2 = W.leb_u(wasm, r)
                if i == code_index:
                    nl, s = W.leb_u(wasm, r2)
                    for _ in range(nl):
                        _, s = W.leb_u(wasm, s)
                        s += 1
                    body = wasm[r2:s] + instrs
                    new_sec = wasm[sec_start:r] + leb_u(len(body)) + body + wasm[r2 + bsize:sec_end]
                    return wasm[:p] + b"\x0a" + leb_u(len(new_sec)) + new_sec + wasm[sec_end:]
                r = r2 + bsize
            raise ValueError("code index out of range")
        p = q + size
    raise ValueError("no code section")


def build(wasm: bytes, body: str = "ops") -> bytes:
    mod = W.Module(wasm)
    b = bytearray(wasm)
    by_name = {f.name: f for f in mod.funcs}
    fn = {f.name: f.idx for f in mod.funcs if f.name}
    mem = mod.memory_image()
    # the constants 0 and 1 of the program's constant table, located through IsZero_79_run (`in != 0`, `1 / in`)
    consts = [imm[0] for _, op, imm in mod.instrs(by_name["IsZero_79_run"]) if op == 0x41 and imm[0] > (1 << 20)]
    zero_addr = next(c for c in consts if W.fr_value(mem, c) == 0)
    one_addr = next(c for c in consts if W.fr_value(mem, c) == 1)
    create, run = by_name["AND_156_create"], by_name["AND_156_run"]
    main_create = by_name["ZkFranchiseProofCircuit_234_create"].idx
    main_run = by_name["ZkFranchiseProofCircuit_234_run"].idx

    def patch_const(fname, old, new, padded=True):
        hits = 0
        for pos, op, imm in mod.instrs(by_name[fname]):
            if op == 0x41 and imm[0] == old:
                n = SW._leb_len(b, pos + 1)
                b[pos + 1:pos + 1 + n] = SW._leb_s_padded(new, n)
                hits += 1
        assert hits == 1, (fname, old, hits)

    def patch_call(fname, old, new):
        hits = 0
        for pos, op, imm in mod.instrs(by_name[fname]):
            if op == 0x10 and imm[0] == old:
                n = SW._leb_len(b, pos + 1)
                b[pos + 1:pos + 1 + n] = SW._leb_u_padded(new, n)
                hits += 1
        assert hits == 1, (fname, old, hits)

    n_in_old = W.getter_const(mod, "getInputSize")
    n_wires_old = W.getter_const(mod, "getWitnessSize")
    patch_const("init", n_in_old, N_IN)
    patch_const("getInputSize", n_in_old, N_IN)
    patch_const("getWitnessSize", n_wires_old, N_WIRES)
    patch_const("AND_156_create", 2, N_IN)                      # inputs the component waits for
    patch_call("init", main_create, create.idx)
    patch_call("setInputSignal", main_run, run.idx)
    for off, data in mod.data:
        at = wasm.find(bytes(data))
        if len(data) == 4096:                                   # input-name hashmap: x[3] at signal 1 + N_OUT
            assert wasm.count(bytes(data)) == 1
            hm = bytearray(4096)
            h = W.fnv1a64("x")
            slot = h & 0xFF
            hm[16 * slot:16 * slot + 16] = struct.pack("<QII", h, 1 + N_OUT, N_IN)
            b[at:at + 4096] = hm
        elif len(data) == 4 * n_wires_old:                      # witness -> signal map: identity
            assert wasm.count(bytes(data)) == 1
            b[at:at + len(data)] = b"".join(struct.pack("<I", i) for i in range(n_wires_old))
    instrs = ops_body(fn, zero_addr, one_addr) if body == "ops" else toint_body(fn)
    return replace_body(bytes(b), run.idx - len(mod.imports), instrs)


if __name__ == "__main__":
    src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(HERE), "artifacts", "zkCensus", "dev", "160", "circuit.wasm")
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/ops_test.wasm"
    open(out, "wb").write(build(open(src, "rb").read()))
    print(out)
