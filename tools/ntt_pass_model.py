#!/usr/bin/env python3
"""Model of the index and twiddle logic of the NTT passes in csrc/ntt.cu over a small NTT-friendly prime: k_ntt_pass
(radix 2, one stage per barrier) and k_ntt_pass4 (radix 4, two stages per barrier), line for line, for every pass
shape the plan produces.  The radix-4 pass was checked against the radix-2 pass here before it ran on a GPU; the GPU
parity tests (tests/test_gpu_kernels.py, test_gpu_sweep.py) tie both to the oracle.

    python tools/ntt_pass_model.py [log2 sizes ...]
"""
import sys
import random
P = 469762049; GEN = 3
TILE_LOG = 11; TILE = 1 << TILE_LOG
def run(logn, lo, hi, DIF, radix4, data, tw):
    n = 1 << logn
    out = list(data)
    rows_log = hi - lo; rows = 1 << rows_log
    g_log = TILE_LOG - rows_log; G = 1 << g_log
    rows_mask = rows - 1; g_mask = G - 1
    col = lo > 0
    for tile in range(n // TILE):
        low0 = 0
        if col:
            lg_log = lo - g_log
            hi_idx = tile >> lg_log; low_grp = tile & ((1 << lg_log) - 1)
            low0 = low_grp * G
            base = (hi_idx << hi) + low0
        else:
            base = tile * TILE
        def addr(i):
            if col:
                g = i & g_mask; k = i >> g_log
                return base + (k << lo) + g
            return base + i
        sm = [out[addr(i)] for i in range(TILE)]
        stride = G if col else 1
        def twd(krow, sb, g):
            j = ((krow & ((1 << sb) - 1)) << lo) + ((low0 + g) if col else 0)
            return tw[j << (logn - 1 - (lo + sb))]
        def stage2(st):
            sb = (rows_log - 1 - st) if DIF else st
            s = lo + sb; d = 1 << sb
            for b in range(TILE // 2):
                if col: g = b & g_mask; kk = b >> g_log
                else: kk = b & (rows_mask >> 1); g = b >> (rows_log - 1)
                k = ((kk >> sb) << (sb + 1)) | (kk & (d - 1))
                i0 = ((k << g_log) + g) if col else ((g << rows_log) + k)
                i1 = i0 + d * stride
                u, v = sm[i0], sm[i1]
                if s == 0:
                    sm[i0] = (u + v) % P; sm[i1] = (u - v) % P; continue
                w = twd(k, sb, g)
                if DIF: sm[i0] = (u + v) % P; sm[i1] = (u - v) * w % P
                else:
                    t = v * w % P; sm[i0] = (u + t) % P; sm[i1] = (u - t) % P
        if not radix4:
            for st in range(rows_log): stage2(st)
        else:
            st = 0
            while st + 1 < rows_log:
                sb0 = (rows_log - 2 - st) if DIF else st; sb1 = sb0 + 1
                d0 = 1 << sb0; unit0 = lo + sb0 == 0
                for q in range(TILE // 4):
                    if col: g = q & g_mask; kk = q >> g_log
                    else: kk = q & (rows_mask >> 2); g = q >> (rows_log - 2)
                    k = ((kk >> sb0) << (sb0 + 2)) | (kk & (d0 - 1))
                    i00 = ((k << g_log) + g) if col else ((g << rows_log) + k)
                    i01 = i00 + d0 * stride; i10 = i00 + 2 * d0 * stride; i11 = i10 + d0 * stride
                    x00, x01, x10, x11 = sm[i00], sm[i01], sm[i10], sm[i11]
                    if DIF:
                        a0 = (x00 + x10) % P; a2 = (x00 - x10) * twd(k, sb1, g) % P
                        a1 = (x01 + x11) % P; a3 = (x01 - x11) * twd(k + d0, sb1, g) % P
                        sm[i00] = (a0 + a1) % P; sm[i10] = (a2 + a3) % P
                        if unit0: sm[i01] = (a0 - a1) % P; sm[i11] = (a2 - a3) % P
                        else:
                            w0 = twd(k, sb0, g); sm[i01] = (a0 - a1) * w0 % P; sm[i11] = (a2 - a3) * w0 % P
                    else:
                        t0, t1 = x01, x11
                        if not unit0:
                            w0 = twd(k, sb0, g); t0 = t0 * w0 % P; t1 = t1 * w0 % P
                        b0 = (x00 + t0) % P; b1 = (x00 - t0) % P; b2 = (x10 + t1) % P; b3 = (x10 - t1) % P
                        u0 = b2 * twd(k, sb1, g) % P; u1 = b3 * twd(k + d0, sb1, g) % P
                        sm[i00] = (b0 + u0) % P; sm[i10] = (b0 - u0) % P; sm[i01] = (b1 + u1) % P; sm[i11] = (b1 - u1) % P
                st += 2
            if st < rows_log: stage2(st)
        for i in range(TILE): out[addr(i)] = sm[i]
    return out
def split_passes(logn):
    lo, hi, top = [], [], logn
    while top > TILE_LOG:
        lo.append(top - 9); hi.append(top); top -= 9
    lo.append(0); hi.append(top)
    return lo, hi
rnd = None


def main(sizes=(11, 12, 13, 14, 15, 17)):
  global rnd
  rnd = random.Random(1)
  for logn in sizes:
      n = 1 << logn
      w = pow(GEN, (P - 1) // n, P)
      tw = [pow(w, e, P) for e in range(n // 2)]
      x = [rnd.randrange(P) for _ in range(n)]
      lo, hi = split_passes(logn)
      for DIF in (True, False):
          a = list(x); b = list(x)
          order = range(len(lo)) if DIF else reversed(range(len(lo)))
          for p in order:
              a = run(logn, lo[p], hi[p], DIF, False, a, tw)
              b = run(logn, lo[p], hi[p], DIF, True, b, tw)
              assert a == b, (logn, DIF, p, lo[p], hi[p])
      print("logn", logn, "passes", list(zip(lo, hi)), "radix-4 == radix-2 (DIF and DIT)")


if __name__ == "__main__":
  main(tuple(int(a) for a in sys.argv[1:]) or (11, 12, 13, 14, 15, 17))
