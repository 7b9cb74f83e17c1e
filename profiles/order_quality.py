import sys, numpy as np
sys.path.insert(0,'/root/repo')
from pico_tree_b200 import datasets as D
tp, q = D.bench_clouds()
lo, hi = tp.min(0), tp.max(0)
def spread(v):
    v = v & 0x3ff
    v = (v | (v << 16)) & 0x030000ff
    v = (v | (v << 8)) & 0x0300f00f
    v = (v | (v << 4)) & 0x030c30c3
    v = (v | (v << 2)) & 0x09249249
    return v
cell = np.clip((q - lo) * (1023.999 / (hi - lo)), 0, 1023).astype(np.uint32)
code = spread(cell[:,0]) | (spread(cell[:,1]) << 1) | (spread(cell[:,2]) << 2)
def quality(order, name):
    qq = q[order]
    n = len(qq) // 32 * 32
    w = qq[:n].reshape(-1, 32, 3)
    ext = (w.max(1) - w.min(1))
    hp = ext.sum(1)
    # distinct fine cells (0.1 m) per warp
    fine = np.floor(w / 0.25).astype(np.int64)
    key = (fine[...,0] * 1000003 + fine[...,1]) * 1000003 + fine[...,2]
    key.sort(axis=1)
    distinct = 1 + (np.diff(key, axis=1) != 0).sum(1)
    print(f"{name:34s} mean half-perimeter of a warp's box {hp.mean():8.3f} m  median {np.median(hp):7.3f}  distinct 0.25 m cells per warp {distinct.mean():6.2f}")
ident = np.arange(len(q))
quality(ident, "input (scan) order")
quality(np.argsort(code >> 14, kind='stable'), "global, top 16 bits (current)")
quality(np.argsort(code >> 6, kind='stable'), "global, top 24 bits")
quality(np.argsort(code, kind='stable'), "global, 30 bits")
for tile in (512, 1024, 2048, 4096, 8192):
    n = len(q)
    pad = (-n) % tile
    c = np.concatenate([code, np.full(pad, 0xffffffff, np.uint32)]).reshape(-1, tile)
    o = np.argsort(c, axis=1, kind='stable') + (np.arange(c.shape[0]) * tile)[:, None]
    o = o.ravel(); o = o[o < n]
    quality(o, f"tile-local 30 bits, tile {tile}")
rng = np.random.default_rng(0)
sh = rng.permutation(len(q))
quality(sh, "shuffled input")
