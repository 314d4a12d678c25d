"""Independent check of the triangle-mesh extension: a float64 numpy Moeller-Trumbore brute force
over all triangles for a set of sampled rays. It shares nothing with the product's plane-first intersector
(csrc/rt_device.cuh tri_hit) or the oracle's restatement of it (oracle/pt_oracle.c hit_tri) except the contract:
closest valid hit with t in [1e-4, 10000], the earlier triangle on ties. TEST INFRASTRUCTURE ONLY."""
import numpy as np

T_MIN, T_MAX = 1e-4, 10000.0


def moller_trumbore(vertices_world, triangles, org, direction, chunk=64):
    """-> (tri index or -1, t, unit geometric normal facing against the ray, distance of the barycentrics from the nearest
    edge) per ray, all float64."""
    v = np.asarray(vertices_world, np.float64)
    tri = np.asarray(triangles, np.int64)
    v0, v1, v2 = v[tri[:, 0]], v[tri[:, 1]], v[tri[:, 2]]
    e1, e2 = v1 - v0, v2 - v0
    nrm = np.cross(e1, e2)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    o = np.asarray(org, np.float64); d = np.asarray(direction, np.float64)
    n = len(o)
    best_t = np.full(n, np.inf); best_i = np.full(n, -1, np.int64); best_edge = np.zeros(n)
    for a in range(0, n, chunk):
        oo, dd = o[a:a + chunk, None, :], d[a:a + chunk, None, :]
        p = np.cross(dd, e2[None])
        det = np.einsum("rtk,tk->rt", p, e1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = oo - v0[None]
            u = np.einsum("rtk,rtk->rt", s, p) * inv
            q = np.cross(s, e1[None])
            w = np.einsum("rtk,rtk->rt", np.broadcast_to(dd, q.shape), q) * inv
            t = np.einsum("rtk,tk->rt", q, e2) * inv
        ok = (np.abs(det) > 1e-14) & (u >= 0) & (w >= 0) & (u + w <= 1) & (t >= T_MIN) & (t <= T_MAX)
        t = np.where(ok, t, np.inf)
        i = np.argmin(t, axis=1)                     # first minimum = earliest triangle on ties
        rows = np.arange(len(i))
        bt = t[rows, i]
        best_t[a:a + chunk] = bt
        best_i[a:a + chunk] = np.where(np.isfinite(bt), i, -1)
        best_edge[a:a + chunk] = np.minimum(np.minimum(u[rows, i], w[rows, i]), 1 - u[rows, i] - w[rows, i])
    hit = best_i >= 0
    gn = np.zeros((n, 3))
    gn[hit] = nrm[best_i[hit]]
    flip = np.einsum("rk,rk->r", gn, d) > 0
    gn[flip] *= -1
    return best_i, np.where(hit, best_t, 0.0), gn, best_edge


def check_against_mt(ids, t, nrm, mt, mesh_id=0, edge_band=1e-4, t_rtol=1e-5):
    """Asserts a tracer's (ids, t, normal) on the same rays agree with moller_trumbore()'s output `mt`:
    ids equal except for rays whose float64 hit lies within `edge_band` (barycentric units) of a triangle edge or of the
    distance limits; t within t_rtol relative; normals parallel. Returns (rays hit, disagreements inside the band)."""
    bi, bt, gn, edge = mt
    hit64 = bi >= 0
    hit = np.asarray(ids) == mesh_id
    differ = hit != hit64
    near = (edge < edge_band) | (np.abs(bt - T_MIN) < 1e-6) | ~hit64
    # a disagreement is only tolerated where the float64 answer itself is marginal: at an edge of the (boundary of the) mesh
    marginal = differ & hit64 & (edge < edge_band)
    # a float32 hit that float64 calls a miss must be a grazing / boundary case too: accept only if it is rare
    phantom = differ & ~hit64
    assert (differ & ~(marginal | phantom)).sum() == 0, "hit/miss disagreement away from any triangle edge"
    assert phantom.sum() <= max(2, int(2e-3 * len(bi))), "too many float32 hits that float64 calls misses: %d" % phantom.sum()
    both = hit & hit64
    assert both.sum() > 0
    rel = np.abs(np.asarray(t, np.float64)[both] - bt[both]) / np.maximum(1.0, np.abs(bt[both]))
    # a ray may cross the crease between two triangles within the band and take the neighbour: same surface, same t
    assert rel.max() <= t_rtol or (rel[~near[both]].max() <= t_rtol and np.quantile(rel, 0.999) <= t_rtol), "t differs: %g" % rel.max()
    interior = both & (edge > edge_band)
    cosang = np.einsum("rk,rk->r", np.asarray(nrm, np.float64)[interior], gn[interior])
    assert cosang.min() > 1 - 1e-5, "normal differs: min cos %g" % cosang.min()
    return int(both.sum()), int(differ.sum())
