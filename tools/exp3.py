"""K1 time vs fraction of non-flat tiles (noise tiles that produce no foreground)."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc
n, h, w = 25, 1024, 1280
st = torch.cuda.current_stream().cuda_stream
outs = [(torch.empty((n, h, w), dtype=torch.uint8, device='cuda'), torch.empty((n, h, w), dtype=torch.int32, device='cuda')) for _ in range(2)]
rng = np.random.default_rng(0)
def make(pheavy, pattern):
    pool = []
    for i in range(8):
        a = np.full((n, h, w), 128, np.uint8)
        noise = rng.integers(-8, 9, size=(n, h, w)).astype(np.int16)
        if pattern == 'random':
            sel = rng.random((n, h // 32, w // 128)) < pheavy
        else:  # clustered: the same central block of every frame
            sel = np.zeros((n, h // 32, w // 128), bool)
            k = int(round(pheavy * 320)); rows = max(1, k // 6)
            sel[:, 16 - rows // 2: 16 - rows // 2 + rows, 2:8] = True
        m = np.repeat(np.repeat(sel, 32, axis=1), 128, axis=2)
        a = np.where(m, (a.astype(np.int16) + noise).astype(np.uint8), a)
        pool.append(torch.from_numpy(a).cuda())
    return pool
for pattern in ('random', 'clustered'):
    for ph in (0.0, 0.1, 0.2, 0.4, 1.0):
        data = make(ph, pattern)
        det = hc.Detector(0, profile=True); det.set_stream(st)
        def step(i): det.enqueue_device(data[i % 8].data_ptr(), n, h, w, 1, None, outs[i & 1][0].data_ptr(), outs[i & 1][1].data_ptr())
        for it in range(6): step(it)
        torch.cuda.synchronize(); det.profile()
        for it in range(30): step(it)
        torch.cuda.synchronize()
        pr = det.profile()
        k1 = pr['preprocess_mask']['ms'] / pr['preprocess_mask']['launches'] * 1e3
        r = det.fetch_results(n)
        print('%-10s heavy %.2f  K1 %.1f us  fg %d' % (pattern, ph, k1, int(r.frames['fg_pixels'].sum())))
        det.close()
