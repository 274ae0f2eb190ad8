"""
Host-side tables of the march PtAP kernels (tg_ptap_march_w, tg_ptap_march in
include/tigar_b200.h) for ONE parametric direction.  Pure numpy -- no device, so
``tests/test_host_logic.py`` can check them on CPU by emulating the kernel's
per-fibre algorithm against a dense M^T A M.

Inputs are the 1-D extraction rows of the direction (what the reference's
per-node loop evaluates, common.py:1497-1509: ``first`` = first spline function
of FE node I, ``vals[I, k]`` = N_{first+k}(x_I), ``m_lo/m_hi`` = the range kept by
the ignore-eps filter, common.py:1508) and the 1-D windows of A_FE (FE x FE) and
of C = M^T A M (IGA x IGA) in that direction.
"""
import numpy as np


def dir_tables(p, pf, first, vals, m_lo, m_hi, loA, hiA, loC, hiC, rmax):
    """Returns a dict of numpy arrays, or None when the direction does not have
    the structure the march kernels assume (the caller then falls back to the
    row-wise kernels):
      mrow[I, k]      eps-filtered M_d[I, first(I)+k]
      tabc[I, q, m]   M_d[loA(I)+q, first(I)+m]   (CTA-tiled kernel), m < p+2
      irec[I]         {lenA(I) | loA(I) << 8, first(I), sbits(I), group(I)}
                      sbits: 2 bits per q, first(loA+q) - first(I) + 1 in {0,1,2}
      SX[I]           exclusive prefix sum of lenA (SX has nfe+1 entries)
      jrec[i]         {loC(i) - (i-p), lenC(i), SY[i] lo32, SY[i] hi32}
      cpad[I]         {0, mrow[I, 0..p], 0, 0}
      grp, gidx       groups of consecutive FE rows sharing first(I) (<= min(rmax, pf) rows)
      GMAX            most window entries of one group, >= 2p+1
    """
    if p > 4:
        return None
    first = np.asarray(first, dtype=np.int64)
    loA, hiA = np.asarray(loA, dtype=np.int64), np.asarray(hiA, dtype=np.int64)
    loC, hiC = np.asarray(loC, dtype=np.int64), np.asarray(hiC, dtype=np.int64)
    nfe, ncp = len(first), len(loC)
    lenA, lenC = hiA - loA + 1, hiC - loC + 1
    TW, TWP = p + 2, (p + 3) & ~1
    j = first[:, None] + np.arange(p + 1)[None, :]
    keep = (j >= np.asarray(m_lo)[:, None]) & (j <= np.asarray(m_hi)[:, None])
    mrow = np.where(keep, vals, 0.0)
    KA = int(lenA.max())
    tabc = np.zeros((nfe, KA, TWP))
    for I in range(nfe):
        for q, J in enumerate(range(loA[I], hiA[I] + 1)):
            m = first[J] - first[I] + np.arange(p + 1)
            k = np.nonzero(mrow[J])[0]
            if k.size == 0:
                continue
            if m[k].min() < 0 or m[k].max() >= TW:
                return None
            tabc[I, q, m[k]] = mrow[J, k]
    # (k = 0, m = p+1) would fall outside the 2p+1 band: must vanish
    if np.any((np.abs(tabc[:, :, p + 1]).sum(axis=1) > 0) & (mrow[:, 0] != 0)):
        return None
    i = np.arange(ncp)
    if np.any(loC < i - p) or np.any(hiC > i + p):
        return None
    if np.any(np.diff(first) < 0):
        return None
    # tables of the warp-task kernel (tg_ptap_march_w)
    SX = np.concatenate([[0], np.cumsum(lenA)]).astype(np.int64)
    SY = np.concatenate([[0], np.cumsum(lenC)]).astype(np.int64)
    sbits = np.zeros(nfe, dtype=np.int64)
    for I in range(nfe):
        for q, J in enumerate(range(loA[I], hiA[I] + 1)):
            sh = first[J] - first[I] + 1
            if not np.any(mrow[J]):
                sh = 1
            if sh < 0 or sh > 2:
                return None
            sbits[I] |= int(sh) << (2 * q)
    if KA > 15 or nfe >= (1 << 23):
        return None
    # groups: runs of FE rows sharing first(I), at most min(rmax, pf) rows each
    grp, gidx = [0], np.zeros(nfe, dtype=np.int64)
    for I in range(1, nfe):
        if first[I] != first[I - 1] or I - grp[-1] >= min(rmax, pf):
            grp.append(I)
        gidx[I] = len(grp) - 1
    grp.append(nfe)
    grp = np.array(grp, dtype=np.int64)
    gsum = np.add.reduceat(lenA, grp[:-1])
    GMAX = int(max(gsum.max(), 2 * p + 1))
    irec = np.stack([lenA | (loA << 8), first, sbits, gidx], axis=1)
    jrec = np.stack([loC - (i - p), lenC, SY[:-1] & 0xffffffff, SY[:-1] >> 32], axis=1)
    cpad = np.zeros((nfe, p + 4))
    cpad[:, 1:p + 2] = mrow
    return dict(mrow=mrow, tabc=tabc, KA=KA, SX=SX, irec=irec, jrec=jrec, cpad=cpad, grp=grp,
                gidx=gidx, GMAX=GMAX)
