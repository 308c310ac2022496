"""Measured dense int8 tensor peak of this GPU: cuBLASLt int8 x int8 -> int32 GEMM (torch._int_mm), burst and sustained.

Measurement only (the roofline denominator of rf_incidence_kernel); nothing in treecl_b200/ imports it.
usage: python tools/ubench/int8_peak.py [out.json]
"""
import json
import sys
import time

import torch


def bench(m, n, k, reps=10):
    a = torch.randint(-2, 2, (m, k), dtype=torch.int8, device='cuda')
    b = torch.randint(-2, 2, (n, k), dtype=torch.int8, device='cuda').t()      # column-major B: the TN layout IMMA kernels want
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch._int_mm(a, b); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~3 s
    t0 = time.time(); n_it = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            torch._int_mm(a, b)
        n_it += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / n_it
    ops = 2.0 * m * n * k
    return {'m': m, 'n': n, 'k': k, 'burst_tops': ops / (best * 1e-3) / 1e12, 'sustained_tops': ops / (sustained * 1e-3) / 1e12,
            'burst_ms': best, 'sustained_ms': sustained}


def main():
    res = {'gpu': torch.cuda.get_device_name(0), 'torch': torch.__version__, 'how': 'torch._int_mm (cuBLASLt int8 -> int32), best of 10 / 3 s loop',
           'shapes': []}
    for (m, n, k) in ((8192, 8192, 8192), (16384, 16384, 4096), (16384, 16384, 3200)):
        try:
            res['shapes'].append(bench(m, n, k))
        except Exception as e:        # noqa
            res['shapes'].append({'m': m, 'n': n, 'k': k, 'error': str(e)[:200]})
    ok = [s for s in res['shapes'] if 'burst_tops' in s]
    if ok:
        res['int8_tops'] = max(s['burst_tops'] for s in ok)
        res['int8_tops_sustained'] = max(s['sustained_tops'] for s in ok)
    line = json.dumps(res)
    print(line)
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(line + '\n')


if __name__ == '__main__':
    main()
