"""Reads off the shared-memory word the tensor core fetches for A element (row m, k) under a layout type / LBO / SBO
(vp_debug_umma_probe): A holds its own word index (in two passes of 10 bits, exact in TF32), B is an 8 x 8 identity in
the K-major 128B-swizzled layout of the forward engine, so D[m][k] = A(m, k).  Usage: python tests/gpu_probe_umma.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_prediction_b200 import lib as L  # noqa: E402

A_BYTES = 64 * 1024


def identity_b():
    b = np.zeros(256, dtype=np.float32)
    for n in range(8):
        for k in range(8):
            if n == k:
                byte = n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4
                b[byte // 4] = 1.0
    return torch.from_numpy(b).cuda()


def fetch_map(a_start, lbo, sbo, layout, mn_major):
    idx = np.arange(A_BYTES // 4)
    out = []
    for code in (idx & 1023, idx >> 10):
        a = torch.from_numpy(code.astype(np.float32)).cuda()
        d = torch.zeros(128, 8, device='cuda')
        L.check(L.lib().vp_debug_umma_probe(L.ptr(a), A_BYTES, L.ptr(identity_b()), 1024, a_start, lbo, sbo, layout, mn_major,
                                            0, 16, 1024, 2, 0, 8, L.ptr(d), L.stream_ptr()))
        torch.cuda.synchronize()
        out.append(d.cpu().numpy().astype(np.int64))
    return (out[0] + 1024 * out[1]) * 4          # byte offsets


def show(title, m):
    print('==', title)
    for r in list(range(0, 12)) + [16, 31, 32, 33, 64, 127]:
        print('  m=%3d:' % r, ' '.join('%6d' % v for v in m[r]))


def swz(L):
    return L ^ (((L >> 7) & 3) << 5)


def check_linear(start, lbo, sbo):
    """32B-atom 128B swizzle: is the fetch address swizzle(start + linear offset), also for starts that are not 128B-aligned?"""
    m = fetch_map(start, lbo, sbo, 1, 1)
    bad = 0
    for r in range(128):
        for k in range(8):
            lin = start + (k % 4) * 128 + (k // 4) * sbo + ((r // 4) % 8) * 16 + (r // 32) * lbo + (r % 4) * 4
            if m[r][k] != swz(lin):
                bad += 1
                if bad <= 4:
                    print('   mismatch m=%d k=%d: fetched %d, model %d (linear %d)' % (r, k, m[r][k], swz(lin), lin))
    print('== 32B-atom MN-major start=%d LBO=%d SBO=%d: %s' % (start, lbo, sbo, 'address-keyed linear model holds' if not bad else '%d mismatches' % bad))
    return m


if __name__ == '__main__':
    for start in (0, 32, 64, 128, 160, 16, 1536 + 96):
        check_linear(start, 1536, 512)
    check_linear(64, 1088, 512)
    if len(sys.argv) < 2 or sys.argv[1] != 'all':
        sys.exit(0)
    show('MN-major, 128B swizzle / 32B atoms (engine wgrad recipe) LBO=8192 SBO=512', fetch_map(0, 8192, 512, 1, 1))
    show('MN-major, no swizzle, LBO=1152 SBO=1152', fetch_map(0, 1152, 1152, 0, 1))
    show('MN-major, no swizzle, LBO=128 SBO=1152', fetch_map(0, 128, 1152, 0, 1))
    show('MN-major, no swizzle, LBO=1152 SBO=128', fetch_map(0, 1152, 128, 0, 1))
    show('MN-major, no swizzle, LBO=1152 SBO=1152, start +16', fetch_map(16, 1152, 1152, 0, 1))
    show('MN-major, 128B swizzle (16B atoms), LBO=1152 SBO=1152', fetch_map(0, 1152, 1152, 2, 1))
    show('K-major, no swizzle, LBO=128 SBO=256', fetch_map(0, 128, 256, 0, 0))
