#!/usr/bin/env python
"""CPU model of the traversal kernel's warp scheduling (no GPU needed): how many lanes does each round serve, and what would
other policies buy?

Step 1 replays the kernel's traversal ORDER on the CPU (numpy, all rays at once): nearest-first BVH4 descent over the scene's
flattened tree with entry-distance culling, leaves of <= 4 triangles, Moller-Trumbore for the hit distance.  Per ray it records the
sequence of events the GPU lane goes through: N (one node step) or L(count, accepted-mask) (one leaf step).
Step 2 feeds those sequences to a model of one persistent warp (32 slots, refill at >= 8 idle slots, majority vote between a
node round and a leaf round — csrc/miro_gpu_api.cu) and to variants of it, and prices the rounds with the SASS instruction
counts of the v11 kernel (profiles/r1_v10_ncu_summary.md): node step 172, triangle iteration 66, Moller-Trumbore block 70, round
bookkeeping 26, ray set-up 103.

usage: tools/sched_sim.py [n_rays] [scene]        (default 131072 incoherent rays of the C2 stand-in, seed as bench.py)
The model knows nothing about latency; it answers "how many instructions / rounds per ray" only.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import bench  # noqa: E402

EMPTY = 0x7fffffff
I_NODE, I_TRI, I_MT, I_ROUND, I_SETUP, I_POP = 172, 66, 70, 26, 103, 15


def flat(sc):
    d = sc.desc()
    if d.n_instances or d.n_mbtris:
        raise SystemExit("sched_sim: static-triangle scenes only (the replay has no instance / motion-blur leaves)")
    nodes = np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_float)), shape=(d.n_nodes, 32)).copy()
    tris = np.ctypeslib.as_array(C.cast(d.tris, C.POINTER(C.c_float)), shape=(d.n_tris, 3, 4)).copy()[:, :, :3]
    return d.root, nodes[:, :24].reshape(-1, 6, 4), nodes[:, 24:28].view(np.int32).copy(), tris


def trace_events(root, bounds, child, tris, rays, max_events=256):
    """Returns events[n, max_events] (0 = none, 1 = node step, 2 + count*16 + accepted_mask = leaf step), n_events[n]."""
    n = len(rays)
    o = rays["o"].astype(np.float32); d = rays["d"].astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d == 0, np.float32(1e12), 1.0 / d).astype(np.float32)
    tmin = rays["tmin"].copy(); best = rays["tmax"].copy()
    stack_ref = np.full((n, 64), EMPTY, np.int64); stack_t = np.zeros((n, 64), np.float32); sp = np.zeros(n, np.int64)
    cur = np.full(n, root, np.int64)
    events = np.zeros((n, max_events), np.int32); ne = np.zeros(n, np.int64)
    alive = np.ones(n, bool)
    it = 0
    while alive.any():
        it += 1
        idx = np.nonzero(alive)[0]
        c = cur[idx]
        is_node = (c >= 0) & (c != EMPTY)
        # ---- node steps
        a = idx[is_node]
        if len(a):
            nd = c[is_node]
            b = bounds[nd]                                   # [m, 6, 4]: lo xyz, hi xyz per child
            t0 = (b[:, 0:3, :] - o[a][:, :, None]) * inv[a][:, :, None]
            t1 = (b[:, 3:6, :] - o[a][:, :, None]) * inv[a][:, :, None]
            tn = np.maximum(np.minimum(t0, t1).max(axis=1), tmin[a][:, None])
            tf = np.minimum(np.maximum(t0, t1).min(axis=1), best[a][:, None])
            ch = child[nd].astype(np.int64)
            hit = (tn <= tf) & (ch != EMPTY)
            key = np.where(hit, tn, np.float32(np.inf))
            order = np.argsort(key, axis=1, kind="stable")
            ks = np.take_along_axis(key, order, 1); cs = np.take_along_axis(ch, order, 1); hs = np.take_along_axis(hit, order, 1)
            nh = hs.sum(axis=1)
            # push far -> near (entries 3, 2, 1 of the sorted order), continue with entry 0
            for j in (3, 2, 1):
                m = hs[:, j]
                rows = a[m]
                stack_ref[rows, sp[rows]] = cs[m, j]; stack_t[rows, sp[rows]] = ks[m, j]; sp[rows] += 1
            cur[a] = np.where(nh > 0, cs[:, 0], EMPTY)
            events[a, ne[a]] = 1; ne[a] += 1
        # ---- leaf steps
        l = idx[~is_node & (c != EMPTY)]
        if len(l):
            u = cur[l] & 0xffffffff
            count = ((u >> 26) & 7) + 1; first = u & ((1 << 26) - 1)
            acc_mask = np.zeros(len(l), np.int64)
            for i in range(int(count.max())):
                m = count > i
                rows = l[m]
                tv = tris[first[m] + i]
                e0 = tv[:, 1] - tv[:, 0]; e1 = tv[:, 2] - tv[:, 0]
                p = np.cross(d[rows], e1); det = (e0 * p).sum(1)
                with np.errstate(divide="ignore", invalid="ignore"):
                    invd = 1.0 / det
                    tvec = o[rows] - tv[:, 0]; q = np.cross(tvec, e0)
                    aa = (tvec * p).sum(1) * invd; bb = (d[rows] * q).sum(1) * invd; tt = (e1 * q).sum(1) * invd
                ok = (det != 0) & (aa >= 0) & (bb >= 0) & (aa + bb <= 1) & (tt >= tmin[rows]) & (tt < best[rows])
                best[rows[ok]] = tt[ok]
                sub = np.nonzero(m)[0][ok]
                acc_mask[sub] |= (1 << i)
            events[l, ne[l]] = 2 + count * 16 + (acc_mask & 15); ne[l] += 1      # (accepted candidates beyond the 4th of a leaf are not priced)
            cur[l] = EMPTY
        # ---- pop (culled by entry distance)
        need = idx[cur[idx] == EMPTY]
        while len(need):
            has = sp[need] > 0
            done = need[~has]; alive[done] = False
            need = need[has]
            if not len(need):
                break
            sp[need] -= 1
            r = stack_ref[need, sp[need]]; t = stack_t[need, sp[need]]
            ok = t < best[need]
            cur[need[ok]] = r[ok]
            need = need[~ok]
        if ne.max() >= max_events - 1:
            raise RuntimeError("event buffer too small")
    return events, ne, best


RESOLVE = 1 << 20      # event code of a deferred Moller-Trumbore evaluation (third phase)


def with_resolve_events(events, ne):
    """Every accepted candidate of a leaf step becomes an event of its own, after the leaf step: the model of a third scheduling
    phase in which the lanes that hold a candidate evaluate it together."""
    n, m = events.shape
    out = np.zeros((n, 2 * m), np.int32); pos = np.zeros(n, np.int64)
    for j in range(int(ne.max())):
        e = events[:, j]
        act = j < ne
        leaf = act & (e >= 2)
        k = np.where(leaf, ((e - 2) & 15), 0)
        acc = ((k & 1) + ((k >> 1) & 1) + ((k >> 2) & 1) + ((k >> 3) & 1))
        rows = np.nonzero(act)[0]
        out[rows, pos[rows]] = np.where(leaf[rows], e[rows] - k[rows], e[rows]); pos[rows] += 1
        for t in range(1, 5):
            r = np.nonzero(acc >= t)[0]
            out[r, pos[r]] = RESOLVE; pos[r] += 1
    return out, pos


def simulate(events, ne, pool=32, refill=8, rays_per_warp=4096, bias=(1, 1), width=32, both_min=0):
    """One persistent warp per block of `rays_per_warp` consecutive rays, all warps simulated at once.  pool: ray slots per warp
    (32 = the kernel; more = a shared-memory pool from which each round picks up to `width` rays of the majority phase)."""
    n = len(ne) // rays_per_warp * rays_per_warp
    W = n // rays_per_warp
    slot_ray = np.full((W, pool), -1, np.int64); slot_pos = np.zeros((W, pool), np.int64)
    nxt = np.zeros(W, np.int64)
    st = dict(node_rounds=0, node_lanes=0, leaf_rounds=0, leaf_lanes=0, tri_iters=0, tri_lane_iters=0, mt_execs=0, mt_lanes=0, refills=0, refill_lanes=0, rounds=0)
    base = np.arange(W) * rays_per_warp
    while True:
        idle = slot_ray < 0
        n_idle = idle.sum(1)
        can = (nxt < rays_per_warp) & (n_idle >= refill)
        if can.any():
            for w in np.nonzero(can)[0]:
                free = np.nonzero(idle[w])[0]
                take = min(len(free), rays_per_warp - nxt[w])
                slot_ray[w, free[:take]] = base[w] + nxt[w] + np.arange(take); slot_pos[w, free[:take]] = 0
                nxt[w] += take
                st["refills"] += 1; st["refill_lanes"] += take
        live = slot_ray >= 0
        if not live.any():
            break
        ev = np.where(live, events[np.maximum(slot_ray, 0), slot_pos], 0)
        at_node = ev == 1; at_leaf = (ev >= 2) & (ev < RESOLVE); at_res = ev == RESOLVE
        n_node = at_node.sum(1); n_leaf = at_leaf.sum(1); n_res = at_res.sum(1)
        res_round = (n_res > n_node) & (n_res > n_leaf)          # third phase: only when it is the largest group
        node_round = ~res_round & (n_node * bias[1] >= n_leaf * bias[0]) & (n_node + n_leaf > 0)
        leaf_round = ~res_round & ~node_round & (n_leaf > 0)
        if both_min > 0:      # a round runs BOTH steps when each phase has at least both_min lanes waiting (node lanes first, then the lanes that were at a leaf)
            both = (n_node >= both_min) & (n_leaf >= both_min)
            node_round |= both; leaf_round |= both
        # pick up to `width` slots of the chosen phase (lowest slot index first)
        chosen = (node_round[:, None] & at_node) | (leaf_round[:, None] & at_leaf) | (res_round[:, None] & at_res)
        if pool > width:
            rank = np.cumsum(chosen, axis=1)
            chosen &= rank <= width
        k = chosen.sum(1)
        st["rounds"] += int((k > 0).sum())
        st["node_rounds"] += int(node_round.sum()); st["node_lanes"] += int((chosen & at_node).sum())
        st["leaf_rounds"] += int(leaf_round.sum()); st["leaf_lanes"] += int((chosen & at_leaf).sum())
        kr = (chosen & at_res).sum(1)
        st["mt_execs"] += int((kr > 0).sum()); st["mt_lanes"] += int(kr.sum())
        lv = np.where(chosen & at_leaf & leaf_round[:, None], ev, 0)
        cnt = np.where(lv >= 2, (lv - 2) >> 4, 0); msk = np.where(lv >= 2, (lv - 2) & 15, 0)
        st["tri_iters"] += int(cnt.max(1).sum()); st["tri_lane_iters"] += int(cnt.sum())
        for i in range(4):
            a = ((msk >> i) & 1).sum(1)
            st["mt_execs"] += int((a > 0).sum()); st["mt_lanes"] += int(a.sum())
        slot_pos[chosen] += 1
        fin = chosen & (slot_pos >= ne[np.maximum(slot_ray, 0)])
        slot_ray[fin] = -1
    st["rays"] = n
    return st


def price(st):
    r = st["rays"]
    instr = (st["node_rounds"] * I_NODE + st["tri_iters"] * I_TRI + st["mt_execs"] * I_MT + st["rounds"] * I_ROUND + st["refills"] * I_SETUP
             + (st["node_rounds"] + st["leaf_rounds"]) * I_POP)
    return dict(instr_per_ray=instr / r, rounds_per_ray=st["rounds"] / r, node_lanes=st["node_lanes"] / max(st["node_rounds"], 1),
                leaf_lanes=st["leaf_lanes"] / max(st["leaf_rounds"], 1), tri_lanes=st["tri_lane_iters"] / max(st["tri_iters"], 1),
                mt_lanes=st["mt_lanes"] / max(st["mt_execs"], 1), refill_lanes=st["refill_lanes"] / max(st["refills"], 1))


def main():
    np.seterr(all="ignore")      # 0 * inf in slab tests of axis-parallel rays, 1 / det of degenerate triangles: masked by the comparisons
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    scene = sys.argv[2] if len(sys.argv) > 2 else "c2_explosion"
    fx = helpers.Fixture(helpers.fixture_path(scene))
    sc = fx.scene()
    root, bounds, child, tris = flat(sc)
    allv = np.concatenate([fx.mesh(k)["vertices"] for k in range(len(fx.names))])
    rays = bench.incoherent_rays(allv.min(0), allv.max(0), n, 0x5EED)
    events, ne, best = trace_events(root, bounds, child, tris, rays)
    nodes = (events == 1).sum() / n; leaves = (events >= 2).sum() / n
    tri = np.where(events >= 2, (events - 2) >> 4, 0).sum() / n
    print("rays %d: %.2f node steps, %.2f leaf steps, %.2f triangle tests per ray, hit fraction %.3f" % (n, nodes, leaves, tri, (best < 1e11).mean()))
    rows = [("kernel as shipped: 32 slots, refill at 8", dict()),
            ("refill at 4", dict(refill=4)), ("refill at 16", dict(refill=16)),
            ("pool of 48 rays, 32 per round", dict(pool=48)), ("pool of 64 rays, 32 per round", dict(pool=64)),
            ("pool of 96 rays, 32 per round", dict(pool=96)),
            ("32 slots + resolve phase", dict(resolve=True)), ("pool of 64 + resolve phase", dict(pool=64, resolve=True)),
            ("both steps per round when each phase has >= 1", dict(both_min=1)), ("... >= 4", dict(both_min=4)),
            ("... >= 8", dict(both_min=8)), ("... >= 12", dict(both_min=12))]
    print("%-44s %9s %9s %6s %6s %6s %6s" % ("policy", "instr/ray", "rounds/ray", "node", "leaf", "tri", "MT"))
    ev3, ne3 = with_resolve_events(events, ne)
    for name, kw in rows:
        kw = dict(kw)
        p = price(simulate(ev3, ne3, **kw) if kw.pop("resolve", False) else simulate(events, ne, **kw))
        print("%-44s %9.1f %9.3f %6.1f %6.1f %6.1f %6.2f" % (name, p["instr_per_ray"], p["rounds_per_ray"], p["node_lanes"], p["leaf_lanes"], p["tri_lanes"], p["mt_lanes"]))


if __name__ == "__main__":
    main()
