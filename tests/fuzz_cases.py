"""Seeded generator of tie-heavy scenes and rays for the differential tests (tests/test_emu_fuzz.py on the CPU
emulator, tests/test_gpu_fuzz.py on the device): exact and jittered cubic / BCC lattices, tight clusters, scenes scaled
by 1e-3 / 1e3, crossed by axis-aligned rays, rays through cell sites and edge midpoints, rays with zero /
denormal-scale direction components, and cameras inside the foam."""
import numpy as np

from radfoam_b200 import foam

SCENE_KINDS = ('lattice', 'bcc', 'cluster', 'scaled', 'uniform')
RAY_KINDS = ('axis', 'vertex', 'random', 'inside', 'zero_comp')


def make_scene(rng, kind, n, deg):
    if kind == 'lattice':
        k = max(3, int(round(n ** (1 / 3))))
        g = np.stack(np.meshgrid(*[np.linspace(-1, 1, k)] * 3, indexing='ij'), -1).reshape(-1, 3)
        pts = (g + rng.normal(0, rng.choice([0.0, 1e-7, 1e-4, 1e-2]), size=g.shape)).astype(np.float32)
    elif kind == 'bcc':
        k = max(3, int(round((n / 2) ** (1 / 3))))
        g = np.stack(np.meshgrid(*[np.arange(k)] * 3, indexing='ij'), -1).reshape(-1, 3).astype(np.float64)
        g = np.concatenate([g, g + 0.5]) / k * 2 - 1
        pts = (g + rng.normal(0, rng.choice([0.0, 1e-6, 1e-3]), size=g.shape)).astype(np.float32)
    elif kind == 'cluster':
        c = rng.normal(size=(5, 3))
        pts = (c[rng.integers(0, 5, n)] + rng.normal(0, 0.05, size=(n, 3)) * rng.choice([1, 10], size=(n, 1))).astype(np.float32)
    elif kind == 'scaled':
        pts = (rng.uniform(-1, 1, size=(n, 3)) * rng.choice([1e-3, 1.0, 1e3])).astype(np.float32)
    else:
        pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    pts = np.unique(pts, axis=0)
    rng.shuffle(pts)
    adj, off = foam.delaunay_adjacency(pts)
    dens = foam.softplus_beta10(rng.normal(rng.choice([-1, 0, 1]), 1.0, size=pts.shape[0])).astype(np.float32) * rng.choice([0.1, 1, 10])
    attrs = foam.make_attributes(rng, pts.shape[0], deg, dens, dc_scale=1.0, sh_sigma=rng.choice([0.0, 0.1, 0.5]))
    return foam.Foam(pts, attrs, adj, off, deg)

def make_rays(rng, f, m):
    ext = float(np.abs(f.points).max())
    kind = rng.choice(['axis', 'vertex', 'random', 'inside', 'zero_comp'])
    if kind == 'axis':
        o = np.zeros((m, 3))
        ax = rng.integers(0, 3, m)
        sg = rng.choice([-1, 1], m)
        d = np.zeros((m, 3))
        d[np.arange(m), ax] = sg
        o = rng.uniform(-ext, ext, size=(m, 3))
        o[np.arange(m), ax] = -sg * 2.5 * ext
        if rng.random() < 0.5:  # through lattice points exactly
            idx = rng.integers(0, f.points.shape[0], m)
            oo = f.points[idx].astype(np.float64).copy()
            oo[np.arange(m), ax] = -sg * 2.5 * ext
            o = oo
    elif kind == 'vertex':
        cam = rng.normal(size=3)
        cam = 3 * ext * cam / np.linalg.norm(cam)
        o = np.tile(cam, (m, 1))
        tgt = f.points[rng.integers(0, f.points.shape[0], m)].astype(np.float64)
        if rng.random() < 0.5:
            tgt = 0.5 * (tgt + f.points[rng.integers(0, f.points.shape[0], m)])
        d = tgt - o
    elif kind == 'inside':
        o = np.tile(f.points[rng.integers(0, f.points.shape[0])].astype(np.float64) + rng.normal(0, 1e-3, 3), (m, 1))
        d = rng.normal(size=(m, 3))
    elif kind == 'zero_comp':
        cam = rng.normal(size=3)
        cam = 3 * ext * cam / np.linalg.norm(cam)
        o = np.tile(cam, (m, 1))
        d = -o + rng.normal(0, 0.3 * ext, size=(m, 3))
        d[np.arange(m), rng.integers(0, 3, m)] = rng.choice([0.0, 1e-30, -1e-25, 1e-12], m)
    else:
        cam = rng.normal(size=3)
        cam = 3 * ext * cam / np.linalg.norm(cam)
        o = np.tile(cam, (m, 1))
        d = -o + rng.normal(0, 0.5 * ext, size=(m, 3))
    d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    rays = np.concatenate([o, d], 1).astype(np.float32)
    start = np.array([foam.nearest_point(f.points, p) for p in np.unique(rays[:, :3], axis=0)])
    uo, inv = np.unique(rays[:, :3], axis=0, return_inverse=True)
    start = start[inv.reshape(-1)].astype(np.uint32)
    return kind, rays, start

def make_case(seed: int):
    """-> (scene kind, ray kind, foam, rays [m,6], start [m], quantiles [m,q] | None, trace settings)."""
    rng = np.random.default_rng(seed)
    scene_kind = rng.choice(SCENE_KINDS)
    deg = int(rng.integers(0, 4))
    f = make_scene(rng, scene_kind, int(rng.integers(40, 400)), deg)
    m = int(rng.choice([1, 7, 33, 100, 257]))
    ray_kind, rays, start = make_rays(rng, f, m)
    q = int(rng.choice([0, 1, 2, 3]))
    dq = None if q == 0 else np.sort(rng.uniform(0, 1, size=(m, q)).astype(np.float32), axis=-1)[:, ::-1].copy()
    kw = {}
    if rng.random() < 0.3:
        kw['max_intersections'] = int(rng.integers(1, 40))
    if rng.random() < 0.3:
        kw['weight_threshold'] = float(rng.choice([0.0, 1e-3, 0.5]))
    return str(scene_kind), str(ray_kind), f, rays, start, dq, kw
