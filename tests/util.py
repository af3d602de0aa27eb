import numpy as np

SEED_DATA, SEED_QUERY, SEED_LEVEL = 0x5EED0001, 0x5EED0002, 0x5EED0003


def uniform_vectors(n, dim, seed):
    """i.i.d. U[0,1) f32, the law of rand_vec (data/functions.rs:2154)."""
    return np.random.default_rng(seed).random((n, dim), dtype=np.float32)


def recall(a_ids, b_ids):
    """mean |a ∩ b| / k over rows, ignoring padding"""
    tot, k = 0.0, a_ids.shape[1]
    for x, y in zip(a_ids, b_ids):
        sx = set(int(v) for v in x if v != 0xFFFFFFFF)
        sy = set(int(v) for v in y if v != 0xFFFFFFFF)
        tot += 1.0 if not sy else len(sx & sy) / len(sy)
    return tot / len(a_ids)


def rmat_edges(scale, edge_factor, seed, a=0.57, b=0.19, c=0.19):
    """RMAT (a,b,c,d) generator, no dedup, self loops kept, ids permuted."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = n * edge_factor
    src = np.zeros(m, np.uint64)
    dst = np.zeros(m, np.uint64)
    for bit in range(scale):
        r = rng.random(m)
        sbit = r >= a + b
        dbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
        src |= sbit.astype(np.uint64) << np.uint64(bit)
        dst |= dbit.astype(np.uint64) << np.uint64(bit)
    perm = rng.permutation(n).astype(np.uint32)
    return n, perm[src.astype(np.int64)], perm[dst.astype(np.int64)]
