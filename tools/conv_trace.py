"""Per-CTA timeline of a conv_halo_kernel launch from the stamps written under FRCNN_CONV_TRACE=<file> (measurement only):
    FRCNN_CONV_TRACE=gpurun_out/conv_trace.bin FRCNN_BENCH_LAYER=conv2_1 FRCNN_BENCH_CFG="128,31" python tools/bench_conv_layers.py 1
    python tools/conv_trace.py gpurun_out/conv_trace.bin
stamps per unit: 0 MMA issuer got the accumulator stage, 1 first operands landed, 2 all MMAs issued, 3 epilogue saw the
accumulator full, 4 epilogue done.  Prints medians of the phases."""
import sys

import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8, 8).astype(np.int64)
start, end = t[:, 7, 7], t[:, 7, 6]
ok = start > 0
t0 = start[ok].min()
print("CTAs %d  span %.1f us  CTA life: median %.1f us  (start median %.1f, end median %.1f)" % (
    ok.sum(), (end[ok].max() - t0) / 1e3, np.median((end - start)[ok]) / 1e3, np.median(start[ok] - t0) / 1e3, np.median(end[ok] - t0) / 1e3))
first = (t[:, 0, 1] - start)[ok & (t[:, 0, 1] > 0)]
print("CTA start -> first operands of unit 0: median %.2f us" % (np.median(first) / 1e3))
rows = []
for u in range(7):
    m = ok & (t[:, u, 0] > 0) & (t[:, u, 4] > 0)
    if m.sum() == 0:
        continue
    mma = (t[:, u, 2] - t[:, u, 1])[m] / 1e3
    wait_ops = (t[:, u, 1] - t[:, u, 0])[m] / 1e3
    full_lag = (t[:, u, 3] - t[:, u, 2])[m] / 1e3
    epi = (t[:, u, 4] - t[:, u, 3])[m] / 1e3
    print("unit %d (%4d CTAs): operands wait %.2f  MMA issue phase %.2f  issue->full %.2f  epilogue %.2f us (medians)" % (
        u, m.sum(), np.median(wait_ops), np.median(mma), np.median(full_lag), np.median(epi)))
    if u > 0:
        m2 = m & (t[:, u - 1, 4] > 0)
        gap = (t[:, u, 0] - t[:, u - 1, 2])[m2] / 1e3
        print("         previous unit's last MMA issue -> this unit's accumulator granted: %.2f us" % np.median(gap))
b = int(np.argmax(np.where(ok, end - start, 0)))
print("longest CTA %d:" % b)
for u in range(7):
    if t[b, u, 0] > 0:
        print("   unit %d: " % u + "  ".join("%.2f" % ((t[b, u, k] - start[b]) / 1e3) for k in range(5)))
