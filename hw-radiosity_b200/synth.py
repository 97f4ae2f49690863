"""Seeded synthetic scenes for the scaling benchmark.

An axis-aligned room of side S whose six walls are m x m grids of quads (two triangles each,
reflectance 101 101 101), k random axis-aligned boxes inside, and one square ceiling light
(255 255 255 / 255 255 255).  The text is the reference's .tri grammar (src/rad.cpp:59-75); the number
of receivers after PrePartitionTriangles is about the total surface area (area threshold 1), so
`leaves` picks S.  The coarse triangle count stays below ~10 k to respect the reference's octree limits.
"""
import numpy as np


def _quad(a, b, c, d, frgb, brgb, out):
    """quad a b c d -> triangles a b c, a c d (the ordering of the reference's .poly/.tri pairs)."""
    for tri in ((a, b, c), (a, c, d)):
        out.append("Triangle \n%d %d %d %d %d %d\n" % (*frgb, *brgb) +
                   "".join("%.6f %.6f %.6f\n" % tuple(p) for p in tri))


def room_scene(leaves, seed=1, wall_grid=20, boxes=40):
    rng = np.random.default_rng(seed)
    # total area ~ 6 S^2 (walls) + boxes; solve for S with the boxes' expected area included afterwards
    S = float(np.sqrt(max(leaves, 1000) / 6.6))
    out = []
    grey, black, white = (101, 101, 101), (0, 0, 0), (255, 255, 255)
    lo, hi = -S / 2, S / 2
    m = wall_grid
    ticks = np.linspace(lo, hi, m + 1)
    light_half = S / 40.0

    def wall(axis, value, flip):
        u, v = [a for a in range(3) if a != axis]
        for i in range(m):
            for j in range(m):
                def P(uu, vv):
                    p = [0.0, 0.0, 0.0]
                    p[axis], p[u], p[v] = value, uu, vv
                    return p
                a, b, c, d = P(ticks[i], ticks[j]), P(ticks[i + 1], ticks[j]), P(ticks[i + 1], ticks[j + 1]), P(ticks[i], ticks[j + 1])
                if flip:
                    a, b, c, d = a, d, c, b
                _quad(a, b, c, d, grey, black, out)
    for axis in range(3):
        wall(axis, lo, False)
        wall(axis, hi, True)
    # ceiling light just below the ceiling (z = hi), facing down
    z = hi - S / 200.0
    _quad([-light_half, -light_half, z], [-light_half, light_half, z], [light_half, light_half, z], [light_half, -light_half, z],
          white, white, out)
    for _ in range(boxes):
        size = rng.uniform(S / 40, S / 12, 3)
        centre = rng.uniform(lo + S / 10, hi - S / 10, 3)
        centre[2] = lo + size[2] / 2 + rng.uniform(0, S / 3)
        a, b = centre - size / 2, centre + size / 2
        x0, y0, z0 = a
        x1, y1, z1 = b
        col = tuple(int(c) for c in rng.integers(60, 200, 3))
        faces = [([x0, y0, z0], [x0, y1, z0], [x1, y1, z0], [x1, y0, z0]), ([x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]),
                 ([x0, y0, z0], [x1, y0, z0], [x1, y0, z1], [x0, y0, z1]), ([x0, y1, z0], [x0, y1, z1], [x1, y1, z1], [x1, y1, z0]),
                 ([x0, y0, z0], [x0, y0, z1], [x0, y1, z1], [x0, y1, z0]), ([x1, y0, z0], [x1, y1, z0], [x1, y1, z1], [x1, y0, z1])]
        for f in faces:
            _quad(*f, col, black, out)
    return "".join(out).encode()
