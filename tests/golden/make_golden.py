"""Generate tests/golden/ref_vectors.npz from the REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
It drives oracle/_ref/liblqro_ref.so (the reference sources compiled unmodified, see
oracle/ref_harness.cpp) and stores inputs + outputs of every stage of the hot path, plus the
reference's own qhull fixture (pointList.txt -> Planes.txt + facetVertices.txt).  The committed
.npz travels to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _libs import REFERENCE_TREE, Reference, RefSim  # noqa: E402


def hexagon(n=6, radius=6.0):
    """start/goal set of the canonical scenario (MAIN:1163-1218): hexagon of radius 6, antipodal goals."""
    d = np.sqrt(3.0)
    pts = np.array([[6.0, 0, 0], [3.0, 3.0 * d, 0], [-3.0, 3.0 * d, 0], [-6.0, 0, 0], [-3.0, -3.0 * d, 0], [3.0, -3.0 * d, 0]])
    goals = pts[[3, 4, 5, 0, 1, 2]]
    return pts[:n].copy(), goals[:n].copy()


def main():
    R = Reference()
    out = {}
    cm = R.control_matrices()
    for k, v in cm.items():
        out["dyn_" + k] = v
    A, B, L, E = cm["A"], cm["B"], cm["L"], cm["E"]
    out["spheres"] = R.spheres()
    out["nominal_input"] = np.array(R.nominal_input())
    # closed-loop powers from the reference's findFG (MAIN:653-662), H = 30
    F, G = np.eye(16), np.zeros((16, 3))
    Fs, Gs = [], []
    for _ in range(30):
        F, G = R.find_fg(A, B, L, E, F, G)
        Fs.append(F.copy())
        Gs.append(G.copy())
    out["F_k"], out["G_k"] = np.array(Fs), np.array(Gs)

    # ---- per-pair cases (H=30 and H=20), outside / inside / too-few-points
    rng = np.random.default_rng(20261018)
    u = R.nominal_input()
    xi_l, xj_l, H_l, st_l, nr_l, dist_l, nrm_l, pl_l, gj_l, pn_l, ridx_l, best_l = ([] for _ in range(12))
    want = {0: 6, 1: 40, 2: 40}
    tries = 0
    while sum(want.values()) > 0 and tries < 20000:
        tries += 1
        xi, xj = np.zeros(16), np.zeros(16)
        xi[12:] = u
        xj[12:] = u
        mode = tries % 3
        d = rng.uniform(0.6, 3.0) if mode == 2 else (rng.uniform(3.0, 14.0) if mode == 1 else rng.uniform(25, 60))
        dirv = rng.normal(size=3)
        dirv /= np.linalg.norm(dirv)
        xi[:3], xj[:3] = dirv * d / 2 + rng.normal(size=3), -dirv * d / 2
        close = rng.uniform(0, 4)
        xi[3:6] = rng.normal(size=3) * 1.0 - dirv * close
        xj[3:6] = rng.normal(size=3) * 1.0 + dirv * close
        xi[6:12] = rng.normal(size=6) * 0.05
        xj[6:12] = rng.normal(size=6) * 0.05
        xi[12:] += rng.normal(size=4) * 0.05
        xj[12:] += rng.normal(size=4) * 0.05
        H = 30 if tries % 2 else 20
        pn = rng.normal(size=3)
        r = R.pair(A, B, L, E, xi, xj, H, 30.0, prev_normal=pn)
        if want[r["status"]] <= 0:
            continue
        want[r["status"]] -= 1
        best = -1
        if r["status"] == 2:
            planes, verts = R.last_hull()
            # nearest facet as the reference scans them (MAIN:825-837)
            reach = r["ellipsoids"][r["reach_idx"]]
            vrel = xi[3:6] - xj[3:6]
            dd = np.abs(np.einsum("fk,fk->f", planes[:, :3], vrel[None, :] - reach[verts[:, 0]]))
            best = int(np.argmin(dd))
        xi_l.append(xi); xj_l.append(xj); H_l.append(H); st_l.append(r["status"]); nr_l.append(r["n_reach"])
        dist_l.append(r["distance"]); nrm_l.append(r["normal"]); pl_l.append(r["plane"]); gj_l.append(r["gjk_stats"])
        pn_l.append(pn); best_l.append(best)
        ri = np.full(1500, -1, dtype=np.int32)
        ri[: r["n_reach"]] = r["reach_idx"]
        ridx_l.append(ri)
    out.update(pair_xi=np.array(xi_l), pair_xj=np.array(xj_l), pair_H=np.array(H_l), pair_status=np.array(st_l),
               pair_nreach=np.array(nr_l), pair_distance=np.array(dist_l), pair_normal=np.array(nrm_l),
               pair_plane=np.array(pl_l), pair_gjk=np.array(gj_l), pair_prev_normal=np.array(pn_l),
               pair_reach_idx=np.array(ridx_l), pair_best_facet=np.array(best_l))

    # ---- LP cases (MAIN:1101-1112), ragged plane counts incl. 0 and infeasible sets
    lp_planes, lp_n, lp_goal, lp_out = [], [], [], []
    for t in range(200):
        n = int(rng.integers(0, 33))
        normals = rng.normal(size=(32, 3))
        normals /= np.linalg.norm(normals, axis=1)[:, None]
        pts = rng.normal(size=(32, 3)) * rng.uniform(0.1, 5)
        if t % 3 == 0:
            pts = -normals * rng.uniform(0, 3, size=(32, 1))
        if t % 7 == 0 and n > 2:
            normals[1] = -normals[0]
            pts[1] = pts[0] + normals[0] * 0.5
        planes = np.concatenate([pts, normals], axis=1).astype(np.float32)
        goal = rng.normal(size=3) * 3
        lp_planes.append(planes); lp_n.append(n); lp_goal.append(goal)
        lp_out.append(R.calculate_new_v(planes[:n], goal))
    out.update(lp_planes=np.array(lp_planes), lp_n=np.array(lp_n), lp_goal=np.array(lp_goal), lp_newv=np.array(lp_out))

    # ---- the reference's own qhull fixture
    out["qh_points"] = np.loadtxt(os.path.join(REFERENCE_TREE, "pointList.txt"), skiprows=2)
    out["qh_planes"] = np.loadtxt(os.path.join(REFERENCE_TREE, "Planes.txt"), skiprows=2)
    out["qh_facets"] = np.loadtxt(os.path.join(REFERENCE_TREE, "facetVertices.txt"), skiprows=1, dtype=np.int64)[:, 1:].astype(np.int32)

    # ---- canonical 6-quad scenario (MAIN:1163-1218), reference loop, MSVC-LCG seed 1
    starts, goals = hexagon()
    sim = RefSim(R, starts, goals, seed=1)
    xs, newvs, sts, pls = [], [], [], []
    for t in range(120):
        x, xt, vg = sim.state()
        newV, planes, status, nplanes, nreach = sim.avoid(30, 30.0)
        xs.append(np.concatenate([x, vg], axis=1)); newvs.append(newV); sts.append(status); pls.append(planes)
        sim.update()
    sim.close()
    out.update(sc6_start=starts, sc6_goal=goals, sc6_x_vgoal=np.array(xs), sc6_newv=np.array(newvs),
               sc6_status=np.array(sts), sc6_planes=np.array(pls))

    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "pairs by status:", np.bincount(np.array(st_l), minlength=3),
          "; 6-quad inside pair-steps:", int((np.array(sts) == 2).sum()), "of", np.array(sts).size)


if __name__ == "__main__":
    main()
