"""Decompose an axis-aligned "union of boxes" triangle mesh into box primitives.

Used by the model compiler for the playroom's concave trimesh colliders (door.obj,
drawer2.obj — reference call sites envs/scenes.py:123,319).  Both meshes are
unions of axis-aligned boxes in mesh space, so an exact
decomposition exists: rasterise onto the grid of unique vertex coordinates,
classify every cell by ray parity, then greedily merge cells into maximal boxes.
"""
import numpy as np


def load_obj(fn):
    V, F = [], []
    for line in open(fn):
        p = line.split()
        if not p:
            continue
        if p[0] == 'v':
            V.append([float(x) for x in p[1:4]])
        elif p[0] == 'f':
            idx = [int(x.split('/')[0]) - 1 for x in p[1:]]
            for k in range(1, len(idx) - 1):
                F.append([idx[0], idx[k], idx[k + 1]])
    return np.array(V, dtype=np.float64), np.array(F, dtype=np.int64)


def _inside(pt, V, F):
    """Generalised winding number of the triangle soup around pt (robust to the
    T-junctions and coincident faces these CAD exports contain)."""
    a = V[F[:, 0]] - pt
    b = V[F[:, 1]] - pt
    c = V[F[:, 2]] - pt
    la = np.linalg.norm(a, axis=1)
    lb = np.linalg.norm(b, axis=1)
    lc = np.linalg.norm(c, axis=1)
    num = np.einsum('ij,ij->i', a, np.cross(b, c))
    den = (la * lb * lc + np.einsum('ij,ij->i', a, b) * lc
           + np.einsum('ij,ij->i', b, c) * la + np.einsum('ij,ij->i', c, a) * lb)
    w = np.sum(2 * np.arctan2(num, den)) / (4 * np.pi)
    return abs(w) > 0.5


def decompose(fn, scale=1.0):
    V, F = load_obj(fn)
    V = V * scale
    axes = [np.unique(np.round(V[:, k], 9)) for k in range(3)]
    nx, ny, nz = [len(a) - 1 for a in axes]
    occ = np.zeros((nx, ny, nz), dtype=bool)
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                c = np.array([(axes[0][i] + axes[0][i + 1]) / 2,
                              (axes[1][j] + axes[1][j + 1]) / 2,
                              (axes[2][k] + axes[2][k + 1]) / 2])
                occ[i, j, k] = _inside(c, V, F)
    # greedy cover: boxes may overlap (harmless in a compound collider); at every
    # step take the largest-volume maximal box that covers a not-yet-covered cell
    import itertools
    boxes = []
    covered = np.zeros_like(occ)
    ext = [np.diff(a) for a in axes]

    def grow(i, j, k, order):
        lo = [i, j, k]
        hi = [i + 1, j + 1, k + 1]
        changed = True
        while changed:
            changed = False
            for ax in order:
                for side in (1, -1):
                    sl = [slice(lo[0], hi[0]), slice(lo[1], hi[1]), slice(lo[2], hi[2])]
                    if side == 1:
                        if hi[ax] >= occ.shape[ax]:
                            continue
                        sl[ax] = hi[ax]
                        if occ[tuple(sl)].all():
                            hi[ax] += 1; changed = True
                    else:
                        if lo[ax] <= 0:
                            continue
                        sl[ax] = lo[ax] - 1
                        if occ[tuple(sl)].all():
                            lo[ax] -= 1; changed = True
        return lo, hi

    cellvol = ext[0][:, None, None] * ext[1][None, :, None] * ext[2][None, None, :]
    total = float((cellvol * occ).sum())
    while (occ & ~covered).any():
        if float((cellvol * (occ & ~covered)).sum()) < 0.01 * total:
            break   # remaining slivers (<1% of the solid volume) are dropped
        best = None
        for (i, j, k) in np.argwhere(occ & ~covered):
            for order in itertools.permutations(range(3)):
                lo, hi = grow(i, j, k, order)
                vol = (ext[0][lo[0]:hi[0]].sum() * ext[1][lo[1]:hi[1]].sum() * ext[2][lo[2]:hi[2]].sum())
                if best is None or vol > best[0] + 1e-15:
                    best = (vol, lo, hi)
        _, lo, hi = best
        covered[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
        l = np.array([axes[0][lo[0]], axes[1][lo[1]], axes[2][lo[2]]])
        h = np.array([axes[0][hi[0]], axes[1][hi[1]], axes[2][hi[2]]])
        boxes.append(((l + h) / 2, (h - l) / 2))
    aabb = (V.min(0), V.max(0))
    return boxes, aabb


if __name__ == '__main__':
    import sys
    for fn, s in [(sys.argv[1] + '/env_meshes/door.obj', 0.0015), (sys.argv[1] + '/env_meshes/drawer2.obj', 1.25)]:
        boxes, aabb = decompose(fn, s)
        print(fn, 'aabb', aabb)
        vol = 0
        for c, h in boxes:
            print('  box c=%s h=%s' % (np.round(c, 5), np.round(h, 5)))
            vol += 8 * h.prod()
        print('  n=%d vol=%g' % (len(boxes), vol))
