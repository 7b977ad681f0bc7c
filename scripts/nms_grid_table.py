"""Markdown table of workloads.NMS20k_grid of a bench.py line.  usage: nms_grid_table.py bench.json"""
import json, sys
g = json.load(open(sys.argv[1]))["workloads"]["NMS20k_grid"]
names = {"aabb2d": "xy rectangle (reference `nms_dim == 2`)", "rot_bev": "rotated BEV", "box3d": "oriented 3-D box (reference `nms_dim == 3`)"}
def fmt(v):
    return "%.0f" % v if v < 10000 else "%.1f ms" % (v * 1e-3)
def k(v):
    return "%d" % v if v < 1000 else "%.1fk" % (v * 1e-3)
print("| L (m) | pair test | iou_thr 1e-5 | 0.1 | 0.5 |\n|---|---|---|---|---|")
for L in ("L40", "L200"):
    for m in ("aabb2d", "rot_bev", "box3d"):
        cells = []
        for i in ("1e-05", "0.1", "0.5"):
            rows = [v for key, v in g.items() if key.startswith("%s_%s_" % (L, m)) and key.endswith("_i" + i)]
            us, kept = [r["us"] for r in rows], [r["kept"] for r in rows]
            cells.append("%s-%s (%s-%s)" % (fmt(min(us)), fmt(max(us)), k(min(kept)), k(max(kept))))
        print("| %s | %s | %s |" % (L[1:], names[m], " | ".join(cells)))
