"""The float64 adjudication of id mismatches (tests/reference_arm.py::adjudicate_mismatches) on constructed cases: it must call a
lost hit a lost hit — on either side — and reserve "edge" / "coincident" for what float32 implementations can legitimately
disagree on.  CPU only."""
import numpy as np

import reference_arm as ra


def geometry():
    # mesh 0: two triangles in the plane z = 0 sharing the diagonal of the unit square; mesh 1: one triangle in z = 0 too (coplanar
    # with mesh 0, covering its first triangle); mesh 2 / 3: a motion-blur pair (z = -1 at time 0, z = -3 at time 1)
    sq = dict(vertices=np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32), vidx=np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    cop = dict(vertices=np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32), vidx=np.array([[0, 1, 2]], np.uint32))
    mb0 = dict(vertices=np.array([[0, 0, -1], [1, 0, -1], [1, 1, -1]], np.float32), vidx=np.array([[0, 1, 2]], np.uint32))
    mb1 = dict(vertices=np.array([[0, 0, -3], [1, 0, -3], [1, 1, -3]], np.float32), vidx=np.array([[0, 1, 2]], np.uint32))
    script = "\n".join(["mesh sq a.obj", "mesh cop b.obj", "mesh mb0 c.obj", "mesh mb1 d.obj", "mbobject mb0 mb1 m", "blas g sq m",
                        "instance g 2 0 0 10  0 2 0 0  0 0 2 0  0 0 0 1"]) + "\n"      # instance 0: the square scaled by 2, moved to x = 10
    return ra.SceneGeometry(script, {"sq": sq, "cop": cop, "mb0": mb0, "mb1": mb1})


def ray(o, d, time=0.0):
    r = np.zeros(1, ra.RAY_DTYPE)
    r["o"] = o; r["d"] = d; r["tmin"] = 1e-3; r["tmax"] = 1e12; r["time"] = time
    return r


def ref_hit(mesh, tri, proxy, t):
    h = np.zeros(1, ra.REFHIT)
    h["mesh"], h["tri"], h["proxy"], h["t"] = mesh, tri, proxy, t
    return h


def classes(adj):
    return {k: v for k, v in adj.items() if k != "hard_idx" and v}


def run(geom, r, product, reference):
    m, t, p = (np.array([x]) for x in product)
    return ra.adjudicate_mismatches(geom, r, [0], m, t, p, reference)


def test_a_lost_hit_is_called_by_its_side():
    g = geometry()
    r = ray([0.7, 0.2, 1.0], [0, 0, -1])          # the middle of triangle 0 of the square, t = 1; the MB triangle lies behind (t = 2)
    # the reference reports the square, the product only the triangle behind it: the product lost a clear, nearer hit
    adj = run(g, r, (2, 0, -1), ref_hit(0, 0, -1, 1.0))
    assert classes(adj) == {"product_missed": 1} and adj["hard_idx"] == [0]
    # the product reports a miss
    adj = run(g, r, (-1, -1, -1), ref_hit(0, 0, -1, 1.0))
    assert classes(adj) == {"product_missed": 1}
    # roles swapped: the reference lost it
    adj = run(g, r, (0, 0, -1), ref_hit(2, 0, -1, 2.0))
    assert classes(adj) == {"reference_missed": 1} and adj["hard_idx"] == []
    adj = run(g, r, (0, 0, -1), ref_hit(-1, -1, -1, -1.0))
    assert classes(adj) == {"reference_missed": 1}


def test_coplanar_surfaces_are_coincident_and_the_diagonal_is_an_edge():
    g = geometry()
    r = ray([0.7, 0.2, 1.0], [0, 0, -1])
    adj = run(g, r, (0, 0, -1), ref_hit(1, 0, -1, 1.0))          # square vs the coplanar triangle over it: same distance
    assert classes(adj) == {"coincident": 1}
    r = ray([0.5, 0.5, 1.0], [0, 0, -1])                         # exactly on the shared diagonal: either triangle of the square
    adj = run(g, r, (0, 0, -1), ref_hit(0, 1, -1, 1.0))
    assert classes(adj) == {"coincident": 1}
    # a product "hit" on a triangle the ray clearly misses, where the reference has a clear hit: unexplained or product_missed, never a tie
    r = ray([0.2, 0.7, 1.0], [0, 0, -1])                         # the middle of triangle 1 of the square
    adj = run(g, r, (0, 0, -1), ref_hit(0, 1, -1, 1.0))
    assert set(classes(adj)) <= {"product_missed", "unexplained"} and adj["hard_idx"] == [0]


def test_edge_tolerance_scales_with_the_distance_of_the_ray_origin():
    g = geometry()
    # 1e-5 beside the diagonal, from 1 unit away: 16 float32 roundings of the translated vertices are 1e-6 — a clear hit, and the
    # product reporting the other triangle is wrong
    r = ray([0.5 + 1e-5, 0.5 - 1e-5, 1.0], [0, 0, -1])
    adj = run(g, r, (0, 1, -1), ref_hit(0, 0, -1, 1.0))
    assert adj["hard_idx"] == [0]
    # the same offset seen from 1000 units away is below the rounding of the translated vertices (6e-5 x 16): an edge case
    r = ray([0.5 + 1e-5, 0.5 - 1e-5, 1000.0], [0, 0, -1])
    adj = run(g, r, (0, 1, -1), ref_hit(0, 0, -1, 1000.0))
    assert classes(adj) in ({"edge": 1}, {"coincident": 1}) and adj["hard_idx"] == []


def test_instances_and_motion_blur_are_resolved_in_object_space():
    g = geometry()
    # instance 0 shows the square scaled by 2 at x = 10..12: a ray down at (11.4, 0.4) hits its triangle 0 (object space 0.7, 0.2)
    r = ray([11.4, 0.4, 5.0], [0, 0, -1])
    t, dist, reach = g.intersect(r[0], 0, 0, 0)
    assert abs(t - 5.0) < 1e-12 and dist > 0.05
    adj = run(g, r, (-1, -1, -1), ref_hit(0, 0, 0, 5.0))
    assert classes(adj) == {"product_missed": 1}
    # the motion-blur triangle sits at z = -1 - 2 * time
    r = ray([0.7, 0.2, 1.0], [0, 0, -1], time=0.5)
    t, dist, reach = g.intersect(r[0], 2, 0, -1)
    assert abs(t - 3.0) < 1e-12
