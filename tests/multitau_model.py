"""CPU model of the formulation used by the warp-per-row multi-tau kernel (csrc/multitau_warp.cu).

Test infrastructure only (never imported by the product).  It restates, with plain Python
integers, the *mathematics* the kernel relies on, so that the reformulation itself can be
checked bit for bit against the oracle (reference corr.cpp:315-431) on the CPU:

  * no per-level compaction: every quantity is a function of the level-0 events (f_i, c_i);
  * G2 of the sparse levels by ONE enumeration of event pairs (i < j): a pair at frame
    distance d contributes to level 0 when d <= 2*dpl and, for d >= 2*dpl, to the two levels
    l0 = bitlength(d) - log2(dpl) - 1 and l0 - 1 only;
  * G2 of the dense levels (L_l <= 4*n0) by bin arrays B_l[t] and windowed products;
  * IP / IF from prefix sums of the counts looked up at frame thresholds;
  * the stale-tail quirk: live counts n_l from the merge levels
    bitlength(f_i ^ f_{i-1}); a conservative trigger; the threshold key K* by replaying the
    boundary walk with rank/select on the level-0 events (V_l[p] = A_m[p], m = max{j<=l: n_j>p}).
"""
import numpy as np

INF = 0x7FFFFFFF


def build_sched(level, tau):
    nl = int(level.max()) + 1 if level.size else 1
    first = [0] * nl
    count = [0] * nl
    lo = [1] * nl
    for i, (l, t) in enumerate(zip(level.tolist(), tau.tolist())):
        loc = t >> l
        if count[l] == 0:
            first[l], lo[l] = i, loc
        count[l] += 1
    return nl, first, count, lo


def bitlength(x):
    return int(x).bit_length()


def select_head(ml, level, p):
    """index of the p-th (0-based) event that starts a bin at `level` (ml_i > level)."""
    k = -1
    for i, m in enumerate(ml):
        if m > level:
            k += 1
            if k == p:
                return i
    raise IndexError


def stale_tail_threshold(n0, n, key_at):
    """Same walk as csrc/multitau.cu stale_tail_threshold, with the final scan replaced by
    the closed form used in the warp kernel (first live key > curmin at position >= first)."""
    first, length = 0, n0
    curmin = INF
    while length > 0:
        half = length >> 1
        mid = first + half
        if mid >= n:
            curmin = min(curmin, key_at(mid))
            length = half
        else:
            if curmin != INF and key_at(mid) > curmin:
                return first, curmin
            first = mid + 1
            length = length - half - 1
    return None


def row_multitau_model(f, c, F, dpl, sched, compat=True, stats=None):
    """-> dict tau_index -> (G2, IP, IF) float32, for one row of integer counts."""
    nl, first, count, lo = sched
    lg = dpl.bit_length() - 1
    assert (1 << lg) == dpl
    n0 = len(f)
    out = {}
    L = [F >> l for l in range(nl)]
    lim = [(F >> l) << l for l in range(nl)]
    # prefix sums and lookups
    ps = np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
    fa = np.asarray(f, np.int64)

    def PS(theta):
        return int(ps[int(np.searchsorted(fa, theta, side="left"))])

    # merge levels and live counts
    ml = [99] + [bitlength(f[i] ^ f[i - 1]) for i in range(1, n0)]
    nlive = []
    for l in range(nl):
        heads = sum(1 for m in ml if m > l)
        dropped = 1 if (n0 > 0 and (f[-1] >> l) >= L[l]) else 0
        nlive.append(heads - dropped)
    # dense levels
    ld = nl
    for l in range(1, nl):
        if L[l] <= 4 * max(n0, 1):
            ld = l
            break
    # compat thresholds
    kstar = [INF] * nl
    if compat and n0 > 0:
        smin = INF
        for l in range(1, nl):
            if nlive[l] < nlive[l - 1]:
                smin = min(smin, f[nlive[l]] >> (l - 1))   # lower bound of A_{l-1}[n_l]
            if nlive[l] < n0 and smin < L[l]:
                n = nlive[l]

                def key_at(p, l=l, n=n):
                    if p < n:
                        return f[select_head(ml, l, p)] >> l
                    lv = l - 1
                    while lv > 0 and nlive[lv] <= p:
                        lv -= 1
                    return f[select_head(ml, lv, p)] >> lv

                if stats is not None:
                    stats["fired"] = stats.get("fired", 0) + 1
                res = stale_tail_threshold(n0, n, key_at)
                if res is not None:
                    fst, curmin = res
                    k1 = key_at(fst)
                    j = int(np.searchsorted(fa, (curmin + 1) << l, side="left"))
                    k2 = (f[j] >> l) if j < n0 else INF
                    kstar[l] = max(k1, k2)
                    if stats is not None:
                        stats["lost"] = stats.get("lost", 0) + 1
    flim = [min(lim[l], kstar[l] << l) if kstar[l] != INF else lim[l] for l in range(nl)]
    klim = [min(L[l], kstar[l]) for l in range(nl)]
    # ---- sparse levels: one pass over event pairs
    W = 2 * dpl + 1
    hist = [[0] * (W + 1) for _ in range(nl)]
    dmax = 17 if ld <= 1 else ((2 * dpl + 1) << (ld - 1))
    top = [lo[l] + count[l] - 1 for l in range(nl)]
    for i in range(n0):
        for j in range(i + 1, n0):
            d = f[j] - f[i]
            if d >= dmax:
                break
            cc = c[i] * c[j]
            if d <= top[0] and d >= lo[0] and count[0] > 0 and f[j] < flim[0]:
                hist[0][d] += cc
            l0 = bitlength(d) - (lg + 1)
            if l0 >= 1 and l0 < ld and count[l0] > 0:
                b = (f[j] >> l0) - (f[i] >> l0)
                if lo[l0] <= b <= top[l0] and f[j] < flim[l0]:
                    hist[l0][b] += cc
            l1 = l0 - 1
            if l1 >= 1 and l1 < ld and count[l1] > 0:
                b = (f[j] >> l1) - (f[i] >> l1)
                if lo[l1] <= b <= top[l1] and f[j] < flim[l1]:
                    hist[l1][b] += cc
    # ---- dense levels: bins
    if ld < nl:
        B = [0] * (L[ld] + 2 * W)
        for i in range(n0):
            if f[i] < lim[ld]:
                B[f[i] >> ld] += c[i]
        for l in range(ld, nl):
            if l > ld:
                nb = [0] * (L[l] + 2 * W)
                for t in range(L[l]):
                    nb[t] = B[2 * t] + B[2 * t + 1]
                B = nb
            for k in range(count[l]):
                tp = lo[l] + k
                s = 0
                for t in range(0, klim[l] - tp):
                    s += B[t] * B[t + tp]
                hist[l][tp] = s
    # ---- outputs
    for l in range(nl):
        total = PS(lim[l])
        s2 = np.float32(2.0 ** (-2 * l))
        s1 = np.float32(2.0 ** (-l))
        for k in range(count[l]):
            tp = lo[l] + k
            neff = L[l] - tp

            def sdiv(num):
                return np.float32(num) / np.float32(neff) if neff > 0 else np.float32(num)

            g2 = sdiv(np.float32(hist[l][tp]) * s2)
            ip = sdiv(np.float32(PS((L[l] - tp) << l)) * s1)
            jf = sdiv(np.float32(total - PS(tp << l)) * s1)
            out[first[l] + k] = (g2, ip, jf)
    return out


def ipif_first_delay(f, F, dpl, sched):
    """Closed forms used by phase 1 of k_multitau_warp: the IF thresholds t' << l ascend with the delay
    index ti and the IP thresholds (L_l - t') << l descend, so an event at frame f only needs
      a = #{ti : t' << l <= f}        (it counts in PS(t' << l) for ti >= a), and
      b = #{ti : (L_l - t') << l > f} (it counts in PS((L_l - t') << l) for ti < b);
    both follow from the bit length of f and of F - f for the regular schedule."""
    nl, first, count, lo = sched
    lg = dpl.bit_length() - 1
    cnt0 = count[0]
    T = sum(count)
    lastl = max([l for l in range(nl) if count[l] > 0] + [0])
    cnt_last = count[lastl] if lastl >= 1 else 0

    def cum(l):  # delays of the levels 1..l
        if l <= 0 or lastl < 1:
            return 0
        return dpl * l if l < lastl else dpl * (lastl - 1) + cnt_last

    if f < 2 * dpl:
        a = min(f, cnt0)
    else:
        ls = f.bit_length() - lg - 1
        a = min(cnt0 + dpl * (ls - 1) + (f >> ls) - dpl, T)
    g = F - f
    lc = g.bit_length() - lg - 1
    b = min(g - 1, cnt0) + cum(lc - 2)
    for l in (lc - 1, lc):
        if l >= 1:
            u = (F >> l) - 1 - (f >> l)
            b += min(max(u - dpl, 0), cum(l) - cum(l - 1))
    return a, min(b, T)
