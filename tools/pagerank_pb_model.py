"""numpy model of the propagation-blocking PageRank engine (cozo_b200/csrc/pagerank.cu, mode 1): the same
layout construction (slot space, hub / group split, group-major entry streams, cell table, bins, row ownership) and
the same four passes (K_A gather, K_B accumulate, K_S straddle, K_F final), written for clarity, not speed.
tests/test_pagerank_model_cpu.py holds it against the oracle for many geometries; the CUDA kernels follow it line by line."""
import numpy as np


def stage(n, src, dst, NH, GS, WIN):
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    od = np.bincount(src, minlength=n)
    order = np.argsort(-od, kind="stable")           # slot -> original id (descending out-degree)
    slot = np.empty(n, np.int64)
    slot[order] = np.arange(n)
    s, d = slot[src], slot[dst]
    o = np.lexsort((s, d))                           # in-CSR in slot space: rows by dst slot, entries by src slot
    s, d = s[o], d[o]
    in_ptr = np.concatenate([[0], np.cumsum(np.bincount(d, minlength=n))])
    od_slot = od[order]
    NH = min(NH, n)
    is_hub = s < NH
    hcnt = np.bincount(d[is_hub], minlength=n)
    mcnt = np.bincount(d[~is_hub], minlength=n)
    hptr = np.concatenate([[0], np.cumsum(hcnt)])
    mptr = np.concatenate([[0], np.cumsum(mcnt)])
    hub_idx = s[is_hub]                              # row-major (entries of a row are sorted, hubs first)
    ms = s[~is_hub] - NH                             # M entries in row-major order == Mpos order
    Mtot = ms.size
    n_src = int((od_slot > 0).sum())
    G = max(0, -(-(n_src - NH) // GS)) if n_src > NH else 0
    NB = -(-Mtot // WIN) if Mtot else 0
    g_of = ms // GS
    perm = np.argsort(g_of, kind="stable")           # stable partition by group keeps Mpos order inside a group
    gcnt = np.bincount(g_of, minlength=max(G, 1))[:max(G, 1)]
    gfirst = np.concatenate([[0], np.cumsum(gcnt)])[:-1]
    gbase = np.zeros(max(G, 1), np.int64)
    pad = 0
    for g in range(G):
        gbase[g] = pad
        pad += (gcnt[g] + 7) // 8 * 8
    a_src = np.zeros(pad + 8, np.int64)
    b_pos = np.zeros(pad + 8, np.int64)
    ctab = np.zeros((NB + 1, max(G, 1)), np.int64)
    for g in range(G):
        ctab[:, g] = gbase[g] + gcnt[g]
    mpos_sorted = perm
    for g in range(G):
        idx = mpos_sorted[gfirst[g]:gfirst[g] + gcnt[g]]            # Mpos values of group g, ascending
        ip = gbase[g] + np.arange(gcnt[g])
        a_src[ip] = ms[idx] % GS
        b_pos[ip] = idx % WIN
        bins = idx // WIN
        for b in range(NB + 1):
            ctab[b, g] = gbase[g] + np.searchsorted(bins, b, side="left")
    rowstart = np.searchsorted(mptr, np.arange(NB + 1) * WIN, side="left")
    rowstart = np.minimum(rowstart, n)
    return dict(n=n, NH=NH, GS=GS, WIN=WIN, G=G, NB=NB, order=order, od=od_slot, hptr=hptr, mptr=mptr,
                hub_idx=hub_idx, a_src=a_src, b_pos=b_pos, ctab=ctab, rowstart=rowstart, gbase=gbase, gcnt=gcnt,
                Mtot=Mtot, pad=pad)


def iterate(st, contrib, scores, base, damping):
    n, NH, GS, WIN, G, NB = (st[k] for k in ("n", "NH", "GS", "WIN", "G", "NB"))
    f32 = np.float32
    val = np.zeros(st["pad"] + 8, f32)
    cpad = np.concatenate([contrib, np.zeros(GS + 8, f32)])
    for g in range(G):                                   # K_A: tile of the group's contributions, entry stream -> val
        tile = cpad[NH + g * GS: NH + (g + 1) * GS]
        lo, hi = st["gbase"][g], st["gbase"][g] + (st["gcnt"][g] + 7) // 8 * 8
        val[lo:hi] = tile[st["a_src"][lo:hi]]
    msum = np.zeros(n, f32)
    part_a = np.zeros(max(NB, 1), f32)
    part_z = np.zeros(max(NB, 1), f32)
    mptr, rowstart = st["mptr"], st["rowstart"]
    for b in range(NB):                                  # K_B
        window = np.full(WIN, np.nan, f32)
        total = 0
        for g in range(G):
            c0, c1 = st["ctab"][b, g], st["ctab"][b + 1, g]
            window[st["b_pos"][c0:c1]] = val[c0:c1]
            total += c1 - c0
        assert total == min(WIN, st["Mtot"] - b * WIN)
        basep = b * WIN
        r0, r1 = rowstart[b], rowstart[b + 1]
        first = mptr[r0]
        cend = min(first, basep + total) - basep if first > basep else 0
        if cend:
            part_a[b] = window[:cend].sum(dtype=f32)
        for r in range(r0, r1):
            s_, e_ = mptr[r] - basep, min(mptr[r + 1], basep + WIN) - basep
            sm = f32(0)
            for j in range(s_, e_):
                sm = f32(sm + window[j])
            if mptr[r + 1] > basep + WIN:
                part_z[b] = sm
            else:
                msum[r] = sm
    for b in range(NB):                                  # K_S
        r0, r1 = rowstart[b], rowstart[b + 1]
        if r1 <= r0:
            continue
        r = r1 - 1
        me = mptr[r + 1]
        if me <= (b + 1) * WIN:
            continue
        sm = part_z[b]
        for bb in range(b + 1, (me - 1) // WIN + 1):
            sm = f32(sm + part_a[bb])
        msum[r] = sm
    hub = contrib[:NH]                                   # K_F
    new = np.empty(n, f32)
    err = 0.0
    for r in range(n):
        hs = f32(0)
        for j in range(st["hptr"][r], st["hptr"][r + 1]):
            hs = f32(hs + hub[st["hub_idx"][j]])
        tot = f32(hs + msum[r])
        new[r] = f32(f32(base) + f32(f32(damping) * tot))
        err += abs(float(new[r]) - float(scores[r]))
    od = st["od"]
    cnew = np.where(od > 0, new / np.maximum(od, 1).astype(f32), f32(0)).astype(f32)
    return new, cnew, err


def pagerank(n, src, dst, damping=0.85, tol=1e-4, max_iter=10, NH=16384, GS=32768, WIN=24576):
    st = stage(n, src, dst, NH, GS, WIN)
    f32 = np.float32
    scores = np.full(n, f32(1.0) / f32(n), f32)
    od = st["od"]
    contrib = np.where(od > 0, scores / np.maximum(od, 1).astype(f32), f32(0)).astype(f32)
    base = (f32(1.0) - f32(damping)) / f32(n)
    it = 0
    while True:
        scores, contrib, err = iterate(st, contrib, scores, base, f32(damping))
        it += 1
        if err < tol or it == max_iter:
            break
    out = np.empty(n, f32)
    out[st["order"]] = scores
    return out, it, err
