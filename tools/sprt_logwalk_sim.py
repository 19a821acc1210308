#!/usr/bin/env python
"""CPU model of the log-domain SPRT tail of ransac_b200/csrc/sprt.cuh (USAC_SPRT_LOGWALK): the same fixed-point logic, step by step,
against the sequential chain of double multiplications (sprt.hpp:205-234; Python floats are IEEE doubles, round to nearest). Every
walk that the fast path decides must give the chain's (good, tested points, tested inliers); the rest is replayed exactly on the device
and is only counted here. Modes: the reference's initial tests for F/E, H and lines, random tests, the C4 pathology (delta > epsilon),
large epsilon; inlier rates at and around the drift-free point.  usage: sprt_logwalk_sim.py [seed=0] [trials=3000]"""
import math
import sys

import numpy as np

FX = 36
FLOOR = -600.0
def exact_chain(flags, lam0, r_in, r_out, A):
    lam = lam0; tin = 0
    for k, f in enumerate(flags):
        ln = lam * (r_in if f else r_out)      # python float = IEEE double, round to nearest
        tin += int(f)
        if ln > A: return False, k + 1, tin
        lam = ln
    return True, len(flags), tin

def fast_chain(flags, lam0, r_in, r_out, A, n_total):
    """returns (status, good, tp, tin); status 'fast' or 'fallback'"""
    if not (1e-200 < lam0 < 1e200 and r_in > 0 and r_out > 0 and A > 0 and math.isfinite(A)): return ('fallback',)
    Lin, Lout, LA, L0 = math.log(r_in), math.log(r_out), math.log(A), math.log(lam0)
    if max(abs(Lin), abs(Lout), abs(LA)) >= 64 or n_total > (1 << 22): return ('fallback',)
    q = lambda x: int(round(x * (1 << FX)))
    Lin_f, Lout_f, LA_f, S = q(Lin), q(Lout), q(LA), q(L0)
    floor_f = q(FLOOR)
    margin = 4 * n_total + 1024
    pos = lambda v: max(v, 0); neg = lambda v: min(v, 0)
    bound_mode = False
    tp = 0; tin = 0
    n = len(flags)
    while tp < n:
        blk = flags[tp:tp + 64]; cnt = len(blk)
        i = int(sum(blk)); o = cnt - i
        up = S + i * pos(Lin_f) + o * pos(Lout_f)
        low = S + i * neg(Lin_f) + o * neg(Lout_f)
        if up < LA_f - margin:
            if not bound_mode and low < floor_f: bound_mode = True
            if bound_mode: S = max(S + i * Lin_f + o * Lout_f, floor_f + i * pos(Lin_f) + o * pos(Lout_f))
            else: S = S + i * Lin_f + o * Lout_f
            tp += cnt; tin += i
            continue
        if bound_mode or low < floor_f: return ('fallback',)
        # per-position
        ik = 0; first_hi = -1; first_band = -1
        for k in range(cnt):
            ik += int(blk[k]); ok = k + 1 - ik
            Sk = S + ik * Lin_f + ok * Lout_f
            if Sk > LA_f + margin and first_hi < 0: first_hi = k
            if abs(Sk - LA_f) <= margin and first_band < 0: first_band = k
        if first_hi < 0 and first_band < 0:
            S = S + i * Lin_f + o * Lout_f; tp += cnt; tin += i; continue
        if first_band >= 0 and (first_hi < 0 or first_band < first_hi): return ('fallback',)
        k = first_hi
        return ('fast', False, tp + k + 1, tin + int(sum(blk[:k + 1])))
    return ('fast', True, tp, tin)

g = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
nfb = 0; nbad = 0; nrej = 0; per_mode = {}
for t in range(trials):
    mode = t % 6
    if mode == 0: eps, delta = 0.2, 0.05
    elif mode == 1: eps, delta = 0.1, 0.01
    elif mode == 2: eps, delta = g.uniform(0.01, 0.9), g.uniform(0.001, 0.5)
    elif mode == 3: eps, delta = 0.004, 0.05          # delta > eps (C4 pathology)
    elif mode == 4: eps, delta = g.uniform(0.3, 0.999), g.uniform(0.0001, 0.01)
    else: eps, delta = 0.001, 0.0001
    r_in = delta / eps; r_out = (1 - delta) / (1 - eps)
    A = float(np.exp(g.uniform(0.1, 12)))
    n = int(g.choice([100, 1000, 10000, 20000]))
    p_in = g.choice([0.0, eps * 0.5, eps, eps * 1.5, delta, (delta + eps) / 2, 0.5, 1.0])
    # around the drift-free point for nastier walks
    if g.random() < 0.3:
        Lin, Lout = math.log(r_in), math.log(r_out)
        if Lin * Lout < 0: p_in = abs(Lout) / (abs(Lin) + abs(Lout))
    flags = (g.random(n) < min(max(p_in, 0), 1)).astype(np.uint8)
    # head: first 64 exact
    good, tp, tin = exact_chain(flags, 1.0, r_in, r_out, A)
    # emulate the handover: exact head of 64
    lam = 1.0; ok = True
    head = min(64, n)
    for k in range(head):
        ln = lam * (r_in if flags[k] else r_out)
        if ln > A: ok = False; break
        lam = ln
    if not ok: continue
    pm = per_mode.setdefault(mode, [0, 0]); pm[0] += 1
    res = fast_chain(flags[head:], lam, r_in, r_out, A, n)
    if res[0] == 'fallback': nfb += 1; pm[1] += 1; continue
    g2, tp2, tin2 = res[1], res[2] + head, res[3] + int(flags[:head].sum())
    if (g2, tp2, tin2) != (good, tp, tin):
        nbad += 1; print("MISMATCH", eps, delta, A, n, p_in, (good, tp, tin), (g2, tp2, tin2))
    nrej += (not good)
print("trials", trials, "replays", nfb, "mismatches", nbad, "rejections decided in the tail", nrej)
names = ["F/E initial (0.2, 0.05)", "H initial (0.1, 0.01)", "random", "delta > epsilon (0.004, 0.05)", "large epsilon", "line initial (0.001, 0.0001)"]
for m_, (tot, fb) in sorted(per_mode.items()):
    print(f"  {names[m_]:32s} walks that reach the tail {tot:5d}, replayed with the exact chain {fb:4d} ({100.0 * fb / max(tot, 1):.1f} %)")
sys.exit(1 if nbad else 0)
