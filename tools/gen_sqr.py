#!/usr/bin/env python3
"""Generator + checker of the 512-bit squaring used by Fp::sqr() on the device (csrc/fp_sqr.inc).

a^2 needs 36 limb products instead of 64: the 28 products a_i a_j (i < j) once, doubled, plus the 8 squares a_i^2.
Layout of the carry chains (the style of Fp::operator*: a `mad.lo.cc / madc.hi.cc` pair per product, which ptxas turns
into one IMAD.WIDE with carry; the CC flag never leaves an asm block):

  * two accumulators by the parity of i + j, so that the products of one row a_i * (a_j, a_j+2, ...) sit on consecutive
    limb pairs and form ONE chain: E (i + j even, limbs 2..13), O (i + j odd, limbs 1..14).  Rows go in increasing i;
    the top of every row's chain is at or above the previous one's, so a carry-out only ever lands on a limb that holds
    earlier carry-outs (no ripple);
  * O += E (13 adds), O <<= 1 in place from the top (15 funnel shifts);
  * the eight squares in one chain with the doubled sum as the addend: T = O + sum a_i^2 2^(64 i).

The script EMULATES the instruction list on random and extreme operands (every carry case: all-ones limbs, p - 1,
single-limb values) and only then writes the asm blocks.  The Montgomery reduction of T reuses the row helpers of
Fp::operator* (fp.cuh, sqr_dev).

    python tools/gen_sqr.py            # check, then rewrite zk_franchise_proof_circuit_b200/csrc/fp_sqr.inc
"""
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M32 = 0xFFFFFFFF


def program():
    """list of blocks; a block is a list of (op, dst, src...) over names a0..a7, E2..E13, O1..O15 and the literal 0"""
    blocks = []
    a = lambda i: f"a{i}"
    E = lambda k: f"E{k}"
    O = lambda k: f"O{k}"

    def row(acc, i, js, first_fresh):
        """products a_i * a_j for j in js (ascending, step 2) accumulated into acc at limbs (i + j, i + j + 1)"""
        blk = []
        for n, j in enumerate(js):
            lo, hi = acc(i + j), acc(i + j + 1)
            blk.append(("mad.lo.cc" if n == 0 else "madc.lo.cc", lo, a(i), a(j), lo))
            blk.append(("madc.hi.cc", hi, a(i), a(j), hi))
        top = acc(i + js[-1] + 2)
        blk.append(("addc", top, top, "0"))
        return blk

    for i in range(7):
        odd = [j for j in range(i + 1, 8) if (i + j) & 1]
        even = [j for j in range(i + 1, 8) if not (i + j) & 1]
        if odd:
            blocks.append(row(O, i, odd, i == 0))
        if even:
            blocks.append(row(E, i, even, i == 0))
    # O += E  (E has limbs 2..13 and a carry limb 14)
    blk = [("add.cc", O(2), O(2), E(2))]
    for k in range(3, 15):
        blk.append(("addc.cc", O(k), O(k), E(k)))
    blk.append(("addc", O(15), O(15), "0"))
    blocks.append(blk)
    # O <<= 1, from the top
    blk = []
    for k in range(15, 1, -1):
        blk.append(("shf.l.wrap", O(k), O(k - 1), O(k), "1"))
    blk.append(("shl", O(1), O(1), "1"))
    blocks.append(blk)
    # T = O + squares (in place); limb 0 is the low half of a0^2
    blk = [("mul.lo", O(0), a(0), a(0)), ("mad.hi.cc", O(1), a(0), a(0), O(1))]
    for i in range(1, 8):
        blk.append(("madc.lo.cc", O(2 * i), a(i), a(i), O(2 * i)))
        blk.append(("madc.hi.cc" if i < 7 else "madc.hi", O(2 * i + 1), a(i), a(i), O(2 * i + 1)))
    blocks.append(blk)
    return blocks


def emulate(blocks, aval):
    reg = {f"a{i}": (aval >> (32 * i)) & M32 for i in range(8)}
    for k in range(16):
        reg[f"E{k}"] = 0
        reg[f"O{k}"] = 0
    get = lambda n: 0 if n == "0" else (1 if n == "1" else reg[n])
    for blk in blocks:
        cc = 0
        for ins in blk:
            op, d = ins[0], ins[1]
            s = [get(x) for x in ins[2:]]
            if op in ("mul.lo", "mul.hi"):
                p = s[0] * s[1]
                reg[d] = (p & M32) if op == "mul.lo" else (p >> 32)
                continue
            if op == "shl":
                reg[d] = (s[0] << s[1]) & M32
                continue
            if op == "shf.l.wrap":
                reg[d] = ((s[1] << s[2]) | (s[0] >> (32 - s[2]))) & M32
                continue
            base = op.split(".")[0]
            cin = cc if base in ("madc", "addc") else 0
            if base in ("mad", "madc"):
                p = s[0] * s[1]
                t = ((p & M32) if ".lo" in op else (p >> 32)) + s[2] + cin
            else:
                t = s[0] + s[1] + cin
            reg[d] = t & M32
            if op.endswith(".cc"):
                cc = t >> 32
            else:
                assert t >> 32 == 0, ("carry lost", ins, hex(aval))
                cc = 0
        # a set CC flag must not be left behind by a block
        assert cc == 0, ("carry leaves the block", blk[-1], hex(aval))
    return sum(reg[f"O{k}"] << (32 * k) for k in range(16))


def check(blocks):
    q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rnd = random.Random(1)
    vals = [0, 1, q - 1, r - 1, (1 << 256) - 1, (1 << 255), M32, M32 << 224, int("f" * 8 + "0" * 8, 16) * ((1 << 256) // ((1 << 64) - 1))]
    vals += [((1 << 256) - 1) ^ (M32 << (32 * i)) for i in range(8)] + [M32 << (32 * i) for i in range(8)]
    vals += [(M32 << (32 * i)) | (M32 << (32 * j)) for i in range(8) for j in range(i)]
    vals += [rnd.getrandbits(256) for _ in range(20000)]
    for v in vals:
        v &= (1 << 256) - 1
        assert emulate(blocks, v) == v * v, hex(v)
    return len(vals)


def emit(blocks):
    cname = lambda n: n if n in ("0", "1") else f"{n[0]}[{n[1:]}]" if n[0] in "EO" else f"a[{n[1:]}]"
    out = ["// generated by tools/gen_sqr.py (checked there by emulation) - do not edit",
           "// in: const uint32_t a[8]; out: uint32_t O[16] = a^2; scratch: uint32_t E[16]; E and O start as zeros"]
    for blk in blocks:
        written, read = [], []
        for ins in blk:
            if ins[1] not in written:
                written.append(ins[1])
        for ins in blk:
            for x in ins[2:]:
                if x not in ("0", "1") and x not in written and x not in read:
                    read.append(x)
        idx = {n: k for k, n in enumerate(written + read)}
        lines = []
        for ins in blk:
            op = ins[0]
            ops = [f"%{idx[x]}" if x in idx else x for x in ins[1:]]
            suffix = ".b32" if op in ("shf.l.wrap", "shl") else ".u32"
            lines.append(f'"{op}{suffix} {", ".join(ops)};\\n\\t"')
        outs = ", ".join(f'"+r"({cname(n)})' for n in written)
        ins_ = ", ".join(f'"r"({cname(n)})' for n in read)
        out.append("asm(" + "\n    ".join(lines) + "\n    : " + outs + "\n    : " + ins_ + ");")
    return "\n".join(out) + "\n"


def main():
    blocks = program()
    n = check(blocks)
    nprod = sum(1 for b in blocks for i in b if ".lo" in i[0] and i[0] != "mul.lo") + 1
    path = os.path.join(ROOT, "zk_franchise_proof_circuit_b200", "csrc", "fp_sqr.inc")
    with open(path, "w") as f:
        f.write(emit(blocks))
    print(f"checked {n} operands, {nprod} limb products, {sum(len(b) for b in blocks)} instructions -> {path}")


if __name__ == "__main__":
    main()
