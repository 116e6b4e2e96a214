"""Reads the per-unit globaltimer stamps conv_head_kernel writes under FRCNN_HEAD_TRACE=<file> (measurement only) and
prints per-CTA timelines: usage  python tools/head_trace.py <file> [n_ctas_to_list]"""
import sys

import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8, 8).astype(np.int64)
G = t.shape[0]
start = t[:, 7, 7]
t0 = start[start > 0].min()
rows = []
for b in range(G):
    end = t[b, 0, 7]
    units = []
    for u in range(7):
        r = t[b, u]
        if r[0] == 0:
            continue
        d = int(r[5])
        units.append(dict(head=d & 255, s=(d >> 8) & 255, nsl=(d >> 16) & 255, tile=d >> 24,
                          prod=(r[6] - t0) / 1e3, mma0=(r[0] - t0) / 1e3, mma1=(r[1] - t0) / 1e3, full=(r[2] - t0) / 1e3,
                          sliced=(r[3] - t0) / 1e3, done=(r[4] - t0) / 1e3 if r[4] else None))
    rows.append((b, (start[b] - t0) / 1e3, (end - t0) / 1e3, units))
ends = np.array([r[2] for r in rows])
print("CTAs %d  kernel span %.1f us  CTA end: mean %.1f  p50 %.1f  max %.1f" % (G, ends.max(), ends.mean(), np.median(ends), ends.max()))
order = np.argsort(-ends)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for b in list(order[:n]) + list(order[-4:]):
    _, s, e, units = rows[b]
    print("cta %3d start %.1f first-operands %.1f end %.1f" % (b, s, (t[b, 7, 6] - t0) / 1e3 if t[b, 7, 6] else -1, e))
    for u in units:
        print("    head %d tile %2d slice %d/%d  prod %.1f  mma %.1f..%.1f  full %.1f  sliced %.1f  done %s" % (
            u["head"], u["tile"], u["s"], u["nsl"], u["prod"], u["mma0"], u["mma1"], u["full"], u["sliced"],
            "%.1f" % u["done"] if u["done"] is not None else "-"))
