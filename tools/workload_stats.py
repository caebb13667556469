"""Developer diagnostic (CPU, uses the oracle as a measuring stick only): shape of the blend workload of a
synthetic scene — list lengths per tile, pairs per pixel, radius distribution.  python tools/workload_stats.py [P]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
case = helpers.scene_case(P, 20241017 + 2, 3)
t = time.time()
o = helpers.run_oracle_cpu(case, backward=False, nthreads=8)
print("oracle fwd %.2fs" % (time.time() - t))
R = o["num_rendered"]; rng = o["ranges"].astype(np.int64); L = rng[:, 1] - rng[:, 0]
print("P", P, "visible", int((o["radii"] > 0).sum()), "R", R, "tiles", len(L))
print("list len percentiles 50/90/99/max:", np.percentile(L, [50, 90, 99]).astype(int), L.max(), "sum", L.sum())
Ls = np.sort(L)[::-1]
print("top 20 lists", Ls[:20])
print("share of instances in lists >= 2048/3072/4096:", [float(L[L >= t].sum()) / R for t in (2048, 3072, 4096)],
      "count", [(L >= t).sum() for t in (2048, 3072, 4096)])
X, Xp, Xa, Xb = (int(v) for v in o["counters"])
print("X evaluated %d  X_p %d  X_a %d  X_b %d ; per px: %.1f %.1f %.1f %.1f" % (X, Xp, Xa, Xb, *(v / (case.width * case.height) for v in (X, Xp, Xa, Xb))))
nc = o["n_contrib"].reshape(case.height, case.width)
print("n_contrib pct 50/90/99/max", np.percentile(nc, [50, 90, 99]).astype(int), nc.max())
r = o["radii"][o["radii"] > 0]
print("radius pct 10/50/90/99/max", np.percentile(r, [10, 50, 90, 99]).astype(int), r.max())
tt = o["tiles_touched"][o["radii"] > 0]
print("tiles touched mean %.2f max %d ; hist" % (tt.mean(), tt.max()), np.bincount(np.minimum(tt, 10)))
# per-tile: depth where the last pixel finishes (max n_contrib) vs list length
gx = (case.width + 15) // 16; gy = (case.height + 15) // 16
pad = np.zeros((gy * 16, gx * 16), np.int64); pad[:case.height, :case.width] = nc
tl = pad.reshape(gy, 16, gx, 16).max(axis=(1, 3)).reshape(-1)
print("sum over tiles of max n_contrib:", tl.sum(), "vs R", R, " ratio", tl.sum() / R)
T = o["final_T"]; print("saturated pixels (T<1e-4 path):", float((T < 1e-3).mean()))
# bbox of {alpha >= 1/255} per instance, clipped to its tile
co = o["conic_opacity"].astype(np.float64); m2 = o["means2D"].astype(np.float64)
a, b, c, op = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
detc = a * c - b * b
covx, covz = c / detc, a / detc
tau = np.log(np.maximum(255.0 * op, 1.0)) + 2e-3
ex, ey = np.sqrt(2 * tau * covx) + 1e-3, np.sqrt(2 * tau * covz) + 1e-3
keys = o["point_list_keys"]; ids = o["point_list"].astype(np.int64)
tile = (keys >> np.uint64(32)).astype(np.int64); tx, ty = tile % gx, tile // gx
x, y = m2[ids, 0], m2[ids, 1]
x0 = np.maximum(tx * 16, np.ceil(x - ex[ids])); x1 = np.minimum(tx * 16 + 15, np.floor(x + ex[ids]))
y0 = np.maximum(ty * 16, np.ceil(y - ey[ids])); y1 = np.minimum(ty * 16 + 15, np.floor(y + ey[ids]))
wd, ht = np.maximum(x1 - x0 + 1, 0), np.maximum(y1 - y0 + 1, 0)
area = (wd * ht).astype(np.int64)
print("clipped bbox area: zero frac %.3f; pct 50/90/95/99/max" % (area == 0).mean(), np.percentile(area, [50, 90, 95, 99]).astype(int), area.max())
for cap in (16, 20, 25, 32, 36, 48, 64):
    print("  area <= %d: %.4f of instances, %.4f of bbox pixels" % (cap, (area <= cap).mean(), area[area <= cap].sum() / area.sum()))
print("sum bbox area", area.sum(), "vs X_b", Xb, " mean area", area.mean())
