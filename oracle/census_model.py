"""CPU restatement of the census circuit: witness + R1CS (TEST INFRASTRUCTURE, not product).

Restates `circuit/census.circom:49-115` (`ZkFranchiseProofCircuit(160)`) and the circomlib 2.0.5
templates it instantiates (`census.circom:3-5`: poseidon.circom, comparators.circom,
smt/smtverifier.circom and what they include: bitify, aliascheck, compconstant, switcher, gates,
smt/smtlevins, smt/smtverifiersm, smt/smtverifierlevel, smt/smthash_poseidon).  circomlib is a
package.json dependency (`circuit/package.json:3`, locked 2.0.5 at
`circuit/package-lock.json:141-143`) and is NOT vendored in the reference, so the template
semantics are restated from its published source; parity is pinned by running the reference's own
`circuit.wasm` (oracle/_ref, built by oracle/wasm2c.py) on the same inputs: every one of the
82,754 wires must be identical (tests/test_oracle_witness.py).

Signals are numbered exactly as circom 2.1.5 numbers them (own signals: outputs, inputs,
intermediates; then sub-components in alphabetical order of instance name, depth-first), so the
wasm's witness->signal table can be applied unchanged.

With emit=True every `<==` / `===` also records an R1CS row over signal ids; `reduce_r1cs` then
eliminates the signals circom's -O2 pass removed (those absent from the wasm's witness table),
giving a constraint system over the wasm's wires.  That system is what the dev proving key is
synthesised from (the reference's real proving_key.zkey / circuit.r1cs are missing blobs,
`.MISSING_LARGE_BLOBS:1-3`).

Pure-Python big-int loops: ~1 s per witness; used on a handful of cases only.
"""
import sys

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]


class AssertFailed(Exception):
    """A `===` constraint does not hold (the wasm raises exceptionHandler(4))."""


class Circuit:
    def __init__(self, poseidon, emit=False):
        self.pos = poseidon          # {'C': {t: [...]}, 'S':..., 'M':..., 'P':...}
        self.V = [1]
        self.emit = emit
        self.lin = []                # list of LC (dict sig->coef) that must equal 0
        self.quad = []               # list of (A, B, C) LCs with A*B = C

    def alloc(self, n=1):
        s = len(self.V)
        self.V.extend([0] * n)
        return s

    # -- linear-combination helpers (dict sig -> coef mod P; sig 0 is the constant 1) ---------
    def ev(self, lc):
        V = self.V
        return sum(c * V[s] for s, c in lc.items()) % P

    def set_lin(self, sig, lc):
        """sig <== lc  (linear)"""
        self.V[sig] = self.ev(lc)
        if self.emit:
            d = dict(lc)
            d[sig] = (d.get(sig, 0) - 1) % P
            self.lin.append(d)

    def set_quad(self, sig, a, b, c=None):
        """sig <== a*b + c"""
        v = self.ev(a) * self.ev(b)
        if c:
            v += self.ev(c)
        self.V[sig] = v % P
        if self.emit:
            cc = {s: (-k) % P for s, k in c.items()} if c else {}
            cc[sig] = (cc.get(sig, 0) + 1) % P
            self.quad.append((dict(a), dict(b), cc))

    def force_lin(self, lc, what=""):
        if self.ev(lc) != 0:
            raise AssertFailed(what)
        if self.emit:
            self.lin.append(dict(lc))

    def force_quad(self, a, b, c, what=""):
        if (self.ev(a) * self.ev(b) - self.ev(c)) % P != 0:
            raise AssertFailed(what)
        if self.emit:
            self.quad.append((dict(a), dict(b), dict(c)))


def L(*terms):
    """L((sig, coef), ...) -> LC"""
    d = {}
    for s, c in terms:
        d[s] = (d.get(s, 0) + c) % P
    return d


# ---------------------------------------------------------------------------------------------
# circomlib templates.  Each class allocates in __init__ (own signals, then children in
# alphabetical order) and computes + constrains in run().
# ---------------------------------------------------------------------------------------------

class Sigma:  # poseidon.circom Sigma: {out, in, in2, in4}
    def __init__(s, c):
        s.c = c
        s.out, s.inp, s.in2, s.in4 = (c.alloc() for _ in range(4))

    def run(s):
        c = s.c
        c.set_quad(s.in2, {s.inp: 1}, {s.inp: 1})
        c.set_quad(s.in4, {s.in2: 1}, {s.in2: 1})
        c.set_quad(s.out, {s.in4: 1}, {s.inp: 1})


class Ark:  # {out[t], in[t]}
    def __init__(s, c, t, C, r):
        s.c, s.t, s.C, s.r = c, t, C, r
        s.out = c.alloc(t)
        s.inp = c.alloc(t)

    def run(s):
        for i in range(s.t):
            s.c.set_lin(s.out + i, {s.inp + i: 1, 0: s.C[i + s.r]})


class Mix:  # {out[t], in[t]}; out[i] = sum_j M[j][i]*in[j]
    def __init__(s, c, t, M):
        s.c, s.t, s.M = c, t, M
        s.out = c.alloc(t)
        s.inp = c.alloc(t)

    def run(s):
        t = s.t
        for i in range(t):
            s.c.set_lin(s.out + i, L(*((s.inp + j, s.M[j * t + i]) for j in range(t))))


class MixLast:  # {out, in[t]}
    def __init__(s, c, t, M, k):
        s.c, s.t, s.M, s.k = c, t, M, k
        s.out = c.alloc()
        s.inp = c.alloc(t)

    def run(s):
        t = s.t
        s.c.set_lin(s.out, L(*((s.inp + j, s.M[j * t + s.k]) for j in range(t))))


class MixS:  # {out[t], in[t]}
    def __init__(s, c, t, S, r):
        s.c, s.t, s.S, s.r = c, t, S, r
        s.out = c.alloc(t)
        s.inp = c.alloc(t)

    def run(s):
        t, S, r = s.t, s.S, s.r
        base = (2 * t - 1) * r
        s.c.set_lin(s.out, L(*((s.inp + i, S[base + i]) for i in range(t))))
        for i in range(1, t):
            s.c.set_lin(s.out + i, L((s.inp + i, 1), (s.inp, S[base + t + i - 1])))


class PoseidonEx:  # {out[nOuts], inputs[nInputs], initialState}
    def __init__(s, c, n_in, n_out):
        s.c = c
        t = s.t = n_in + 1
        s.n_out = n_out
        s.RF, s.RP = 8, N_ROUNDS_P[t - 2]
        C, S, M, Pm = (c.pos[k][t] for k in "CSMP")
        s.C = C
        s.out = c.alloc(n_out)
        s.inputs = c.alloc(n_in)
        s.initial = c.alloc()
        RF, RP = s.RF, s.RP
        # children alphabetical: ark, mix, mixLast, mixS, sigmaF, sigmaP
        s.ark = []
        for k in range(RF):
            if k <= RF // 2:
                off = k * t
            else:
                off = (RF // 2 + 1) * t + RP + (k - RF // 2 - 1) * t
            s.ark.append(Ark(c, t, C, off))
        s.mix = [Mix(c, t, Pm if k == RF // 2 - 1 else M) for k in range(RF - 1)]
        s.mixLast = [MixLast(c, t, M, k) for k in range(n_out)]
        s.mixS = [MixS(c, t, S, r) for r in range(RP)]
        s.sigmaF = [[Sigma(c) for _ in range(t)] for _ in range(RF)]
        s.sigmaP = [Sigma(c) for _ in range(RP)]

    def run(s):
        c, t, RF, RP = s.c, s.t, s.RF, s.RP
        h = RF // 2
        wire = c.set_lin
        for j in range(t):
            wire(s.ark[0].inp + j, {(s.inputs + j - 1) if j > 0 else s.initial: 1})
        s.ark[0].run()
        for r in range(h - 1):
            for j in range(t):
                src = s.ark[0].out + j if r == 0 else s.mix[r - 1].out + j
                wire(s.sigmaF[r][j].inp, {src: 1})
                s.sigmaF[r][j].run()
            for j in range(t):
                wire(s.ark[r + 1].inp + j, {s.sigmaF[r][j].out: 1})
            s.ark[r + 1].run()
            for j in range(t):
                wire(s.mix[r].inp + j, {s.ark[r + 1].out + j: 1})
            s.mix[r].run()
        for j in range(t):
            wire(s.sigmaF[h - 1][j].inp, {s.mix[h - 2].out + j: 1})
            s.sigmaF[h - 1][j].run()
        for j in range(t):
            wire(s.ark[h].inp + j, {s.sigmaF[h - 1][j].out: 1})
        s.ark[h].run()
        for j in range(t):
            wire(s.mix[h - 1].inp + j, {s.ark[h].out + j: 1})
        s.mix[h - 1].run()
        for r in range(RP):
            src = s.mix[h - 1].out if r == 0 else s.mixS[r - 1].out
            wire(s.sigmaP[r].inp, {src: 1})
            s.sigmaP[r].run()
            wire(s.mixS[r].inp, {s.sigmaP[r].out: 1, 0: s.C[(h + 1) * t + r]})
            for j in range(1, t):
                wire(s.mixS[r].inp + j, {src + j: 1})
            s.mixS[r].run()
        for r in range(h - 1):
            for j in range(t):
                src = s.mixS[RP - 1].out + j if r == 0 else s.mix[h + r - 1].out + j
                wire(s.sigmaF[h + r][j].inp, {src: 1})
                s.sigmaF[h + r][j].run()
            for j in range(t):
                wire(s.ark[h + r + 1].inp + j, {s.sigmaF[h + r][j].out: 1})
            s.ark[h + r + 1].run()
            for j in range(t):
                wire(s.mix[h + r].inp + j, {s.ark[h + r + 1].out + j: 1})
            s.mix[h + r].run()
        for j in range(t):
            wire(s.sigmaF[RF - 1][j].inp, {s.mix[RF - 2].out + j: 1})
            s.sigmaF[RF - 1][j].run()
        for i in range(s.n_out):
            for j in range(t):
                wire(s.mixLast[i].inp + j, {s.sigmaF[RF - 1][j].out: 1})
            s.mixLast[i].run()
            wire(s.out + i, {s.mixLast[i].out: 1})


class Poseidon:  # {out, inputs[n]}; child pEx
    def __init__(s, c, n):
        s.c, s.n = c, n
        s.out = c.alloc()
        s.inputs = c.alloc(n)
        s.pEx = PoseidonEx(c, n, 1)

    def run(s):
        c = s.c
        c.set_lin(s.pEx.initial, {})
        for i in range(s.n):
            c.set_lin(s.pEx.inputs + i, {s.inputs + i: 1})
        s.pEx.run()
        c.set_lin(s.out, {s.pEx.out: 1})


class IsZero:  # comparators.circom: {out, in, inv}
    def __init__(s, c):
        s.c = c
        s.out, s.inp, s.inv = c.alloc(), c.alloc(), c.alloc()

    def run(s):
        c = s.c
        v = c.V[s.inp]
        c.V[s.inv] = pow(v, P - 2, P) if v else 0          # inv <-- in!=0 ? 1/in : 0
        c.set_quad(s.out, {s.inp: P - 1}, {s.inv: 1}, {0: 1})
        c.force_quad({s.inp: 1}, {s.out: 1}, {}, "IsZero")


class IsEqual:  # {out, in[2]}; child isz
    def __init__(s, c):
        s.c = c
        s.out = c.alloc()
        s.inp = c.alloc(2)
        s.isz = IsZero(c)

    def run(s):
        s.c.set_lin(s.isz.inp, {s.inp + 1: 1, s.inp: P - 1})
        s.isz.run()
        s.c.set_lin(s.out, {s.isz.out: 1})


class ForceEqualIfEnabled:  # {enabled, in[2]}; child isz
    def __init__(s, c):
        s.c = c
        s.enabled = c.alloc()
        s.inp = c.alloc(2)
        s.isz = IsZero(c)

    def run(s):
        s.c.set_lin(s.isz.inp, {s.inp + 1: 1, s.inp: P - 1})
        s.isz.run()
        s.c.force_quad({0: 1, s.isz.out: P - 1}, {s.enabled: 1}, {}, "ForceEqualIfEnabled")


class Num2Bits:  # bitify.circom: {out[n], in}
    def __init__(s, c, n):
        s.c, s.n = c, n
        s.out = c.alloc(n)
        s.inp = c.alloc()

    def run(s):
        c = s.c
        v = c.V[s.inp]
        lc = {}
        for i in range(s.n):
            c.V[s.out + i] = (v >> i) & 1                   # out[i] <-- (in >> i) & 1
            c.force_quad({s.out + i: 1}, {s.out + i: 1, 0: P - 1}, {}, "Num2Bits bit")
            lc[s.out + i] = (1 << i) % P
        lc[s.inp] = P - 1
        c.force_lin(lc, "Num2Bits sum")


class CompConstant:  # compconstant.circom: {out, in[254], parts[127], sout}; child num2bits
    def __init__(s, c, ct):
        s.c, s.ct = c, ct
        s.out = c.alloc()
        s.inp = c.alloc(254)
        s.parts = c.alloc(127)
        s.sout = c.alloc()
        s.num2bits = Num2Bits(c, 135)

    def run(s):
        c, ct = s.c, s.ct
        b = (1 << 128) - 1
        a = 1
        e = 1
        tot = {}
        for i in range(127):
            clsb = (ct >> (2 * i)) & 1
            cmsb = (ct >> (2 * i + 1)) & 1
            slsb, smsb = s.inp + 2 * i, s.inp + 2 * i + 1
            if cmsb == 0 and clsb == 0:
                c.set_quad(s.parts + i, {smsb: (-b) % P}, {slsb: 1}, L((smsb, b), (slsb, b)))
            elif cmsb == 0 and clsb == 1:
                c.set_quad(s.parts + i, {smsb: a % P}, {slsb: 1},
                           L((slsb, -a), (smsb, b), (smsb, -a), (0, a)))
            elif cmsb == 1 and clsb == 0:
                c.set_quad(s.parts + i, {smsb: b % P}, {slsb: 1}, L((smsb, -a), (0, a)))
            else:
                c.set_quad(s.parts + i, {smsb: (-a) % P}, {slsb: 1}, {0: a % P})
            tot[s.parts + i] = 1
            b -= e
            a += e
            e *= 2
        c.set_lin(s.sout, tot)
        c.set_lin(s.num2bits.inp, {s.sout: 1})
        s.num2bits.run()
        c.set_lin(s.out, {s.num2bits.out + 127: 1})


class AliasCheck:  # aliascheck.circom: {in[254]}; child compConstant
    def __init__(s, c):
        s.c = c
        s.inp = c.alloc(254)
        s.compConstant = CompConstant(c, P - 1)

    def run(s):
        c = s.c
        for i in range(254):
            c.set_lin(s.compConstant.inp + i, {s.inp + i: 1})
        s.compConstant.run()
        c.force_lin({s.compConstant.out: 1}, "AliasCheck")


class Num2BitsStrict:  # bitify.circom Num2Bits_strict: {out[254], in}; children aliasCheck, n2b
    def __init__(s, c):
        s.c = c
        s.out = c.alloc(254)
        s.inp = c.alloc()
        s.aliasCheck = AliasCheck(c)
        s.n2b = Num2Bits(c, 254)

    def run(s):
        c = s.c
        c.set_lin(s.n2b.inp, {s.inp: 1})
        s.n2b.run()
        for i in range(254):
            c.set_lin(s.out + i, {s.n2b.out + i: 1})
            c.set_lin(s.aliasCheck.inp + i, {s.n2b.out + i: 1})
        s.aliasCheck.run()


class LessThan:  # comparators.circom: {out, in[2]}; child n2b
    def __init__(s, c, n):
        s.c, s.n = c, n
        s.out = c.alloc()
        s.inp = c.alloc(2)
        s.n2b = Num2Bits(c, n + 1)

    def run(s):
        c = s.c
        c.set_lin(s.n2b.inp, {s.inp: 1, 0: (1 << s.n) % P, s.inp + 1: P - 1})
        s.n2b.run()
        c.set_lin(s.out, {0: 1, s.n2b.out + s.n: P - 1})


class LessEqThan:  # {out, in[2]}; child lt
    def __init__(s, c, n):
        s.c = c
        s.out = c.alloc()
        s.inp = c.alloc(2)
        s.lt = LessThan(c, n)

    def run(s):
        c = s.c
        c.set_lin(s.lt.inp, {s.inp: 1})
        c.set_lin(s.lt.inp + 1, {s.inp + 1: 1, 0: 1})
        s.lt.run()
        c.set_lin(s.out, {s.lt.out: 1})


class AND:  # gates.circom: {out, a, b}
    def __init__(s, c):
        s.c = c
        s.out, s.a, s.b = c.alloc(), c.alloc(), c.alloc()

    def run(s):
        s.c.set_quad(s.out, {s.a: 1}, {s.b: 1})


class MultiAND:  # gates.circom: {out, in[n]}; children and1 | and2, ands[0], ands[1]
    def __init__(s, c, n):
        s.c, s.n = c, n
        s.out = c.alloc()
        s.inp = c.alloc(n)
        if n == 2:
            s.and1 = AND(c)
        elif n > 2:
            s.and2 = AND(c)
            s.ands = [MultiAND(c, n // 2), MultiAND(c, n - n // 2)]

    def run(s):
        c, n = s.c, s.n
        if n == 1:
            c.set_lin(s.out, {s.inp: 1})
        elif n == 2:
            c.set_lin(s.and1.a, {s.inp: 1})
            c.set_lin(s.and1.b, {s.inp + 1: 1})
            s.and1.run()
            c.set_lin(s.out, {s.and1.out: 1})
        else:
            n1 = n // 2
            for i in range(n1):
                c.set_lin(s.ands[0].inp + i, {s.inp + i: 1})
            for i in range(n - n1):
                c.set_lin(s.ands[1].inp + i, {s.inp + n1 + i: 1})
            s.ands[0].run()
            s.ands[1].run()
            c.set_lin(s.and2.a, {s.ands[0].out: 1})
            c.set_lin(s.and2.b, {s.ands[1].out: 1})
            s.and2.run()
            c.set_lin(s.out, {s.and2.out: 1})


class Switcher:  # switcher.circom: {outL, outR, sel, L, R, aux}
    def __init__(s, c):
        s.c = c
        s.outL, s.outR, s.sel, s.L, s.R, s.aux = (c.alloc() for _ in range(6))

    def run(s):
        c = s.c
        c.set_quad(s.aux, {s.R: 1, s.L: P - 1}, {s.sel: 1})
        c.set_lin(s.outL, {s.aux: 1, s.L: 1})
        c.set_lin(s.outR, {s.aux: P - 1, s.R: 1})


class SMTHash1:  # smthash_poseidon.circom: {out, key, value}; child h = Poseidon(3)
    def __init__(s, c):
        s.c = c
        s.out, s.key, s.value = c.alloc(), c.alloc(), c.alloc()
        s.h = Poseidon(c, 3)

    def run(s):
        c = s.c
        c.set_lin(s.h.inputs, {s.key: 1})
        c.set_lin(s.h.inputs + 1, {s.value: 1})
        c.set_lin(s.h.inputs + 2, {0: 1})
        s.h.run()
        c.set_lin(s.out, {s.h.out: 1})


class SMTHash2:  # {out, L, R}; child h = Poseidon(2)
    def __init__(s, c):
        s.c = c
        s.out, s.L, s.R = c.alloc(), c.alloc(), c.alloc()
        s.h = Poseidon(c, 2)

    def run(s):
        c = s.c
        c.set_lin(s.h.inputs, {s.L: 1})
        c.set_lin(s.h.inputs + 1, {s.R: 1})
        s.h.run()
        c.set_lin(s.out, {s.h.out: 1})


class SMTLevIns:  # smtlevins.circom: {levIns[n], enabled, siblings[n], done[n-1]}; isZero[n]
    def __init__(s, c, n):
        s.c, s.n = c, n
        s.levIns = c.alloc(n)
        s.enabled = c.alloc()
        s.siblings = c.alloc(n)
        s.done = c.alloc(n - 1)
        s.isZero = [IsZero(c) for _ in range(n)]

    def run(s):
        c, n = s.c, s.n
        for i in range(n):
            c.set_lin(s.isZero[i].inp, {s.siblings + i: 1})
            s.isZero[i].run()
        c.force_quad({s.isZero[n - 1].out: 1, 0: P - 1}, {s.enabled: 1}, {}, "SMTLevIns last")
        c.set_lin(s.levIns + n - 1, {0: 1, s.isZero[n - 2].out: P - 1})
        c.set_lin(s.done + n - 2, {s.levIns + n - 1: 1})
        for i in range(n - 2, 0, -1):
            c.set_quad(s.levIns + i, {0: 1, s.done + i: P - 1}, {0: 1, s.isZero[i - 1].out: P - 1})
            c.set_lin(s.done + i - 1, {s.levIns + i: 1, s.done + i: 1})
        c.set_lin(s.levIns, {0: 1, s.done: P - 1})


class SMTVerifierSM:  # smtverifiersm.circom
    NAMES = ("st_top st_i0 st_iold st_inew st_na is0 levIns fnc prev_top prev_i0 prev_iold "
             "prev_inew prev_na prev_top_lev_ins prev_top_lev_ins_fnc").split()

    def __init__(s, c):
        s.c = c
        for nm in s.NAMES:
            setattr(s, nm, c.alloc())

    def run(s):
        c = s.c
        c.set_quad(s.prev_top_lev_ins, {s.prev_top: 1}, {s.levIns: 1})
        c.set_quad(s.prev_top_lev_ins_fnc, {s.prev_top_lev_ins: 1}, {s.fnc: 1})
        c.set_lin(s.st_top, {s.prev_top: 1, s.prev_top_lev_ins: P - 1})
        c.set_lin(s.st_inew, {s.prev_top_lev_ins: 1, s.prev_top_lev_ins_fnc: P - 1})
        c.set_quad(s.st_iold, {s.prev_top_lev_ins_fnc: 1}, {0: 1, s.is0: P - 1})
        c.set_quad(s.st_i0, {s.prev_top_lev_ins: 1}, {s.is0: 1})
        c.set_lin(s.st_na, {s.prev_na: 1, s.prev_inew: 1, s.prev_iold: 1, s.prev_i0: 1})


class SMTVerifierLevel:  # smtverifierlevel.circom; children proofHash, switcher
    NAMES = "root st_top st_i0 st_iold st_inew st_na sibling old1leaf new1leaf lrbit child".split()

    def __init__(s, c):
        s.c = c
        for nm in s.NAMES:
            setattr(s, nm, c.alloc())
        s.aux = c.alloc(2)
        s.proofHash = SMTHash2(c)
        s.switcher = Switcher(c)

    def run(s):
        c = s.c
        c.set_lin(s.switcher.L, {s.child: 1})
        c.set_lin(s.switcher.R, {s.sibling: 1})
        c.set_lin(s.switcher.sel, {s.lrbit: 1})
        s.switcher.run()
        c.set_lin(s.proofHash.L, {s.switcher.outL: 1})
        c.set_lin(s.proofHash.R, {s.switcher.outR: 1})
        s.proofHash.run()
        c.set_quad(s.aux, {s.proofHash.out: 1}, {s.st_top: 1})
        c.set_quad(s.aux + 1, {s.old1leaf: 1}, {s.st_iold: 1})
        c.set_quad(s.root, {s.new1leaf: 1}, {s.st_inew: 1}, {s.aux: 1, s.aux + 1: 1})


class SMTVerifier:  # smtverifier.circom
    def __init__(s, c, n):
        s.c, s.n = c, n
        s.enabled, s.root = c.alloc(), c.alloc()
        s.siblings = c.alloc(n)
        s.oldKey, s.oldValue, s.isOld0, s.key, s.value, s.fnc = (c.alloc() for _ in range(6))
        # children alphabetical
        s.areKeyEquals = IsEqual(c)
        s.checkRoot = ForceEqualIfEnabled(c)
        s.hash1New = SMTHash1(c)
        s.hash1Old = SMTHash1(c)
        s.keysOk = MultiAND(c, 4)
        s.levels = [SMTVerifierLevel(c) for _ in range(n)]
        s.n2bNew = Num2BitsStrict(c)
        s.n2bOld = Num2BitsStrict(c)
        s.sm = [SMTVerifierSM(c) for _ in range(n)]
        s.smtLevIns = SMTLevIns(c, n)

    def run(s):
        c, n = s.c, s.n
        w = c.set_lin
        w(s.hash1Old.key, {s.oldKey: 1})
        w(s.hash1Old.value, {s.oldValue: 1})
        s.hash1Old.run()
        w(s.hash1New.key, {s.key: 1})
        w(s.hash1New.value, {s.value: 1})
        s.hash1New.run()
        w(s.n2bOld.inp, {s.oldKey: 1})
        s.n2bOld.run()
        w(s.n2bNew.inp, {s.key: 1})
        s.n2bNew.run()
        for i in range(n):
            w(s.smtLevIns.siblings + i, {s.siblings + i: 1})
        w(s.smtLevIns.enabled, {s.enabled: 1})
        s.smtLevIns.run()
        for i in range(n):
            m = s.sm[i]
            if i == 0:
                w(m.prev_top, {s.enabled: 1})
                w(m.prev_i0, {})
                w(m.prev_inew, {})
                w(m.prev_iold, {})
                w(m.prev_na, {0: 1, s.enabled: P - 1})
            else:
                q = s.sm[i - 1]
                w(m.prev_top, {q.st_top: 1})
                w(m.prev_i0, {q.st_i0: 1})
                w(m.prev_inew, {q.st_inew: 1})
                w(m.prev_iold, {q.st_iold: 1})
                w(m.prev_na, {q.st_na: 1})
            w(m.is0, {s.isOld0: 1})
            w(m.fnc, {s.fnc: 1})
            w(m.levIns, {s.smtLevIns.levIns + i: 1})
            m.run()
        last = s.sm[n - 1]
        c.force_lin({last.st_na: 1, last.st_iold: 1, last.st_inew: 1, last.st_i0: 1, 0: P - 1},
                    "SMTVerifier final state")
        for i in range(n - 1, -1, -1):
            lv, m = s.levels[i], s.sm[i]
            w(lv.st_top, {m.st_top: 1})
            w(lv.st_i0, {m.st_i0: 1})
            w(lv.st_iold, {m.st_iold: 1})
            w(lv.st_inew, {m.st_inew: 1})
            w(lv.st_na, {m.st_na: 1})
            w(lv.sibling, {s.siblings + i: 1})
            w(lv.old1leaf, {s.hash1Old.out: 1})
            w(lv.new1leaf, {s.hash1New.out: 1})
            w(lv.lrbit, {s.n2bNew.out + i: 1})
            if i == n - 1:
                w(lv.child, {})
            else:
                w(lv.child, {s.levels[i + 1].root: 1})
            lv.run()
        w(s.areKeyEquals.inp, {s.oldKey: 1})
        w(s.areKeyEquals.inp + 1, {s.key: 1})
        s.areKeyEquals.run()
        w(s.keysOk.inp, {s.fnc: 1})
        w(s.keysOk.inp + 1, {0: 1, s.isOld0: P - 1})
        w(s.keysOk.inp + 2, {s.areKeyEquals.out: 1})
        w(s.keysOk.inp + 3, {s.enabled: 1})
        s.keysOk.run()
        c.force_lin({s.keysOk.out: 1}, "SMTVerifier keysOk")
        w(s.checkRoot.enabled, {s.enabled: 1})
        w(s.checkRoot.inp, {s.levels[0].root: 1})
        w(s.checkRoot.inp + 1, {s.root: 1})
        s.checkRoot.run()


INPUT_ORDER = [  # main's own signals: public inputs first (circuit-compiler.sh:85-88), then private
    ("electionId", 2), ("nullifier", 1), ("voteHash", 2), ("sikRoot", 1), ("censusRoot", 1),
    ("voteWeight", 1), ("availableWeight", 1), ("address", 1), ("password", 1), ("signature", 1),
    ("censusSiblings", None), ("sikSiblings", None)]


class ZkFranchiseProofCircuit:  # circuit/census.circom:49-115
    def __init__(s, c, n_levels=160):
        s.c = c
        real = s.real = n_levels + 1                                   # census.circom:50
        s.sig = {}
        for nm, k in INPUT_ORDER:
            s.sig[nm] = c.alloc(k if k else real)
        s.censusVerifier = SMTVerifier(c, real)                        # :92
        s.checkNullifier = ForceEqualIfEnabled(c)                      # :111
        s.checkWeight = LessEqThan(c, 252)                             # :69
        s.computedNullifier = Poseidon(c, 4)                           # :105
        s.sik = Poseidon(c, 3)                                         # :74
        s.sikVerifier = SMTVerifier(c, real)                           # :79

    def run(s, inputs):
        c, g, real = s.c, s.sig, s.real
        for nm, k in INPUT_ORDER:
            v = inputs[nm]
            v = v if isinstance(v, list) else [v]
            assert len(v) == (k if k else real), nm
            for i, x in enumerate(v):
                c.V[g[nm] + i] = int(x) % P
        w = c.set_lin
        cw = s.checkWeight                                             # :69-72
        w(cw.inp, {g["voteWeight"]: 1})
        w(cw.inp + 1, {g["availableWeight"]: 1})
        cw.run()
        c.force_lin({cw.out: 1, 0: P - 1}, "checkWeight")
        w(s.sik.inputs, {g["address"]: 1})                             # :74-77
        w(s.sik.inputs + 1, {g["password"]: 1})
        w(s.sik.inputs + 2, {g["signature"]: 1})
        s.sik.run()
        for ver, root, sib, val in ((s.sikVerifier, "sikRoot", "sikSiblings", s.sik.out),
                                    (s.censusVerifier, "censusRoot", "censusSiblings",
                                     g["availableWeight"])):         # :79-103
            w(ver.enabled, {0: 1})
            w(ver.fnc, {})
            w(ver.root, {g[root]: 1})
            for i in range(real):
                w(ver.siblings + i, {g[sib] + i: 1})
            w(ver.oldKey, {})
            w(ver.oldValue, {})
            w(ver.isOld0, {})
            w(ver.key, {g["address"]: 1})
            w(ver.value, {val: 1})
            ver.run()
        cn = s.computedNullifier                                       # :105-109
        w(cn.inputs, {g["signature"]: 1})
        w(cn.inputs + 1, {g["password"]: 1})
        w(cn.inputs + 2, {g["electionId"]: 1})
        w(cn.inputs + 3, {g["electionId"] + 1: 1})
        cn.run()
        ck = s.checkNullifier                                          # :111-114
        w(ck.enabled, {0: 1})
        w(ck.inp, {cn.out: 1})
        w(ck.inp + 1, {g["nullifier"]: 1})
        ck.run()


def build(poseidon, inputs, emit=False, n_levels=160):
    """Run the model; returns the Circuit (c.V = all circom signals)."""
    sys.setrecursionlimit(10000)
    c = Circuit(poseidon, emit=emit)
    z = ZkFranchiseProofCircuit(c, n_levels)
    z.run(inputs)
    return c


def witness(poseidon, wmap, inputs):
    """82,754-wire witness (list of ints) in wasm order."""
    c = build(poseidon, inputs)
    V = c.V
    return [V[s] for s in wmap]


# ---------------------------------------------------------------------------------------------
# R1CS reduction onto the wasm's wires
# ---------------------------------------------------------------------------------------------

def reduce_r1cs(c: Circuit, wmap, verbose=False):
    """Eliminate every signal that is not a wasm wire (what circom -O2 did), return rows over wires.

    Returns (rows, n_eliminated) with rows = [(A, B, C)] of {wire: coef} dicts, A*B = C.
    """
    kept = set(wmap)
    sub = {}

    def resolve(lc):
        """Rewrite lc until it mentions no substituted signal (with path compression)."""
        out = {}
        stack = [(s, k) for s, k in lc.items()]
        while stack:
            s, k = stack.pop()
            e = sub.get(s)
            if e is None:
                v = (out.get(s, 0) + k) % P
                if v:
                    out[s] = v
                else:
                    out.pop(s, None)
                continue
            if any(t in sub for t in e):
                e = resolve(e)
                sub[s] = e
            for t, kk in e.items():
                v = (out.get(t, 0) + k * kk) % P
                if v:
                    out[t] = v
                else:
                    out.pop(t, None)
        return out

    def add_linear(lc):
        lc = resolve(lc)
        cand = [s for s in lc if s not in kept and s != 0]
        if not cand:
            if lc:
                raise RuntimeError(f"linear row over kept wires only: {lc}")
            return 0
        s = max(cand)
        inv = pow(lc[s], P - 2, P)
        sub[s] = {t: (-k * inv) % P for t, k in lc.items() if t != s}
        return 1

    trivial = 0
    for lc in c.lin:
        if not add_linear(lc):
            trivial += 1
    quad = c.quad
    rnd = 0
    while True:
        rnd += 1
        nxt = []
        new = 0
        for A, B, C in quad:
            A, B, C = resolve(A), resolve(B), resolve(C)
            ca = all(s == 0 for s in A)
            cb = all(s == 0 for s in B)
            if ca or cb:
                # (const)*B = C  ->  linear
                if ca:
                    k = A.get(0, 0)
                    lin = {s: (k * v) % P for s, v in B.items()}
                else:
                    k = B.get(0, 0)
                    lin = {s: (k * v) % P for s, v in A.items()}
                for s, v in C.items():
                    lin[s] = (lin.get(s, 0) - v) % P
                lin = {s: v for s, v in lin.items() if v}
                if add_linear(lin):
                    new += 1
                else:
                    trivial += 1
            else:
                nxt.append((A, B, C))
        quad = nxt
        if verbose:
            print(f"round {rnd}: {len(quad)} quadratic rows, {new} new eliminations, "
                  f"{len(sub)} eliminated, {trivial} trivial", file=sys.stderr)
        if new == 0:
            break
    wire = {s: i for i, s in enumerate(wmap)}
    rows = []
    for A, B, C in quad:
        rows.append(tuple({wire[s]: v for s, v in X.items()} for X in (A, B, C)))
    return rows, len(sub)
