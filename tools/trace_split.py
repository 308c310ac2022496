"""Summarise the per-tile clock stamps written by a -DPS_TRACE build (TREEDIST_TRACE=path)."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
names = ['zero+sync', 'events', 'k-loop', 'barrier', 'out']
d = np.diff(t[:, :6], axis=1)
print('tiles', len(t), 'events/tile mean', t[:, 7].mean())
for k, n in enumerate(names):
    print('%-10s mean %8.0f  p50 %8.0f  p90 %8.0f cycles' % (n, d[:, k].mean(), np.percentile(d[:, k], 50), np.percentile(d[:, k], 90)))
tot = t[:, 5] - t[:, 0]
print('total      mean %8.0f  p50 %8.0f  p90 %8.0f cycles' % (tot.mean(), np.percentile(tot, 50), np.percentile(tot, 90)))
# concurrency per SM: span / sum
for sm in range(3):
    m = t[t[:, 6] == sm]
    if len(m):
        span = m[:, 5].max() - m[:, 0].min()
        print('sm', sm, 'tiles', len(m), 'span', span, 'sum of CTA lifetimes', (m[:, 5] - m[:, 0]).sum(), 'avg concurrency %.2f' % ((m[:, 5] - m[:, 0]).sum() / span))
