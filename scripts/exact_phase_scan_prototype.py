"""Prototype (numpy, CPU) of a PARALLEL and BIT-EXACT version of WORLD's running phase

    total[i] = fl(total[i-1] + inc[i]),   inc[i] = 2 pi f0(i) / fs > 0        (synthesis.cpp GetTimeBase)

which csrc/synth.cu phase_scan_kernel evaluates as one dependent fp64 addition chain per utterance (1.6 ms for a 6.5 s
utterance, whatever the batch size).  Floating-point addition is not associative, so an ordinary parallel prefix sum changes
the roundings -- and pulse positions are threshold decisions on fmod(total, 2 pi).  But the rounding has structure:

  while total stays inside one binade [2^e, 2^(e+1)) it is an integer multiple T of u = 2^(e-52), and
      fl(T u + x) = (T + X + d) u,   X = floor(x / u),  rho = x - X u,
      d = 0 if rho < u/2,  1 if rho > u/2,  and for an exact tie rho == u/2 (round half to even)  d = (T + X) & 1.

  So one addition is the integer map T -> T + X + d(parity of T): a segment of additions is described by TWO integers
  (the total increment for an even and for an odd incoming T), and these pairs compose associatively:
      (A then B).even = A.even + B[parity(A.even)],   (A then B).odd = A.odd + B[parity(1 + A.odd)].
  A parallel prefix over the pairs therefore reproduces every sequentially rounded total exactly.  The binade changes only
  ~log2(total) times per utterance (total grows monotonically); each crossing is one ordinary addition, after which the scan
  restarts with u doubled.

`exact_scan` below implements this with blocked composition (the block structure a GPU kernel would use) and is checked against
the sequential float64 loop in tests/test_oracle_synthesis.py, including the 500 Hz / 16 kHz case whose increments tie exactly.
This file is a design aid for the next round; the product path still uses the sequential kernel."""
import math

import numpy as np


def sequential(inc):
    out = np.empty_like(inc)
    t = 0.0
    for i, x in enumerate(inc):
        t = t + float(x)
        out[i] = t
    return out


def _maps(x, u):
    """Per-element integer maps for the unit u: (even, odd) increments as Python ints."""
    X = np.floor(x / u)                       # exact: u is a power of two
    rho = x - X * u                           # exact
    up = rho > u / 2
    tie = rho == u / 2
    Xi = X.astype(np.int64)
    base = Xi + up.astype(np.int64)
    even = base + (tie & ((Xi & 1) == 1)).astype(np.int64)   # incoming T even: T + X odd  <=> X odd  -> round up to even
    odd = base + (tie & ((Xi & 1) == 0)).astype(np.int64)    # incoming T odd:  T + X odd  <=> X even -> round up
    return even, odd


def _compose(a, b):
    ae, ao = a
    be, bo = b
    return (ae + (be if ae % 2 == 0 else bo), ao + (be if (1 + ao) % 2 == 0 else bo))


def exact_scan(inc, block=256):
    """Bit-exact running sum of positive float64 increments by blocked composition of two-state integer maps."""
    inc = np.asarray(inc, np.float64)
    n = len(inc)
    out = np.empty(n)
    i = 0
    t = 0.0
    while i < n:
        if t == 0.0 or not math.isfinite(t):
            t = t + float(inc[i])             # 0 + x is exact; start of the utterance
            out[i] = t
            i += 1
            continue
        e = math.frexp(t)[1] - 1              # t in [2^e, 2^(e+1))
        u = math.ldexp(1.0, e - 52)
        T0 = int(t / u)                       # exact integer in [2^52, 2^53)
        limit = 1 << 53                       # T must stay below: next binade
        even, odd = _maps(inc[i:], u)
        # blocked prefix: compose the maps of every block, scan the block summaries, then expand inside the blocks --
        # exactly the three phases of a GPU scan.  (numpy loops stand in for the parallel steps.)
        m = len(even)
        nb = (m + block - 1) // block
        summaries = []
        for b in range(nb):
            acc = (0, 0)
            for k in range(b * block, min(m, (b + 1) * block)):
                acc = _compose(acc, (int(even[k]), int(odd[k])))
            summaries.append(acc)
        T = T0
        done = 0
        crossed = False
        for b in range(nb):
            T_after = T + (summaries[b][0] if T % 2 == 0 else summaries[b][1])
            lo, hi = b * block, min(m, (b + 1) * block)
            if T_after < limit:                # the whole block stays in the binade: expand it
                Tk = T
                for k in range(lo, hi):
                    Tk = Tk + (int(even[k]) if Tk % 2 == 0 else int(odd[k]))
                    out[i + k] = Tk * u
                T = T_after
                done = hi
                continue
            Tk = T                             # the block contains the crossing: expand up to it
            for k in range(lo, hi):
                Tn = Tk + (int(even[k]) if Tk % 2 == 0 else int(odd[k]))
                if Tn >= limit:
                    crossed = True
                    break
                Tk = Tn
                out[i + k] = Tk * u
                done = k + 1
            T = Tk
            break
        t = T * u
        i += done
        if crossed and i < n:                  # the crossing addition itself: one ordinary rounded add, new binade
            t = t + float(inc[i])
            out[i] = t
            i += 1
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for fs, f0 in ((16000, 500.0), (22050, 500.0), (48000, 500.0)):
        inc = np.full(40000, 2.0 * math.pi * f0 / fs)
        inc[5000:9000] = 2.0 * math.pi * rng.uniform(80, 300, 4000) / fs
        a, b = sequential(inc), exact_scan(inc)
        print(fs, "bit-exact:", np.array_equal(a, b))
