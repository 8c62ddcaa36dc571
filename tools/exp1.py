"""Experiment driver: K1 static vs dynamic schedule, per-frame CCL durations, phase stamps, steady-state step time."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
n, h, w = 25, 1024, 1280
pool = [torch.from_numpy(synth.bottle_batch(n, h, w, start_index=1000 * i)).cuda() for i in range(8)]
st = torch.cuda.current_stream().cuda_stream

def k1_time(tag):
    det = hc.Detector(0, profile=True, phase_timing=True)
    det.set_stream(st)
    for it in range(4): det.detect_device(pool[it % 8].data_ptr(), n, h, w)
    det.profile()
    for it in range(20): det.enqueue_device(pool[it % 8].data_ptr(), n, h, w)
    torch.cuda.synchronize()
    pr = det.profile()
    k1 = det.phase_times()[200:206]
    print(tag, 'K1 us %.1f  CCL us %.1f' % (pr['preprocess_mask']['ms'] / pr['preprocess_mask']['launches'] * 1e3,
          pr['ccl_frame_fused']['ms'] / max(pr['ccl_frame_fused']['launches'], 1) * 1e3),
          '| flat wait %.2f proc %.2f n=%d | nonflat wait %.2f proc %.2f n=%d' % (k1[0]/max(k1[2],1)/1e3, k1[1]/max(k1[2],1)/1e3, k1[2], k1[3]/max(k1[5],1)/1e3, k1[4]/max(k1[5],1)/1e3, k1[5]))
    q = det.phase_times()[248:256]; print('   refill us/tile %.2f' % (q[6] / 8000 / 1e3))
    print('   K1 CTA lifetimes: kernel span %.1f us, mean CTA life %.1f, max CTA life %.1f, earliest end at %.1f, latest start at %.1f' % (
        (q[1] - q[0]) / 1e3, q[2] / 592 / 1e3, q[3] / 1e3, (q[4] - q[0]) / 1e3, (q[5] - q[0]) / 1e3))
    return det

det = k1_time('static ')
pt = det.phase_times()
per = [(pt[208 + f] >> 32, (pt[208 + f] >> 16) & 0xffff, pt[208 + f] & 0xffff) for f in range(n)]
print('per-frame CCL (us, nw, ncomp):', [(round(a / 1e3, 1), b, c) for a, b, c in per])
slow = int(np.argmax([a for a, _, _ in per]))
det.close()
os.environ['HV_K1_DYNAMIC'] = '1'
det = k1_time('dynamic')
det.close()
os.environ['HV_PHASE_FRAME'] = str(slow)
det = hc.Detector(0, profile=True, phase_timing=True); det.set_stream(st)
for it in range(3): det.detect_device(pool[0].data_ptr(), n, h, w)
pt = np.array(det.phase_times()[:192], dtype=np.int64).reshape(12, 16); t0 = pt[0].min()
print('   warp1 iteration start cycles', det.phase_times()[196:203]); q2 = det.phase_times()[193:196]; print('   max steps thread: find steps', (q2[0] >> 32) & 0xffff, 'union iters', q2[0] >> 48, 'tid', q2[0] & 0xffff, '| totals: find steps', q2[1], 'union iters', q2[2]); q = det.phase_times()[192]; print('stamps of frame', slow, 'P2c slowest thread: cycles', q >> 32, 'tid', (q >> 20) & 0xfff, 'vert unions', (q >> 8) & 0xff, 'left unions', (q >> 16) & 0xf, 'segs', q & 0xff)
for i in range(12):
    if pt[i].max() > 0: print('  stamp', i, 'warps min/max us %.2f %.2f' % ((pt[i].min() - t0) / 1e3, (pt[i].max() - t0) / 1e3), ' per warp:', ' '.join('%.1f' % ((x - t0) / 1e3) for x in pt[i]))
det.close()
for env in ({}, {'HV_NO_PDL': '1'}):
    for k in ('HV_NO_PDL',): os.environ.pop(k, None)
    os.environ.update(env)
    det2 = hc.Detector(0); det2.set_stream(st)
    outs = [(torch.empty((n, h, w), dtype=torch.uint8, device='cuda'), torch.empty((n, h, w), dtype=torch.int32, device='cuda')) for _ in range(2)]
    def step(i): det2.enqueue_device(pool[i % 8].data_ptr(), n, h, w, 1, None, outs[i & 1][0].data_ptr(), outs[i & 1][1].data_ptr())
    for it in range(20): step(it)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record()
    for it in range(K): step(it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(env, 'us/step %.1f frames/s %.0f  pipeline frac %.3f' % (ms * 1e3, n / ms * 1e3, n * h * w * 6 / ms / 1e6 / 6546.6))
    det2.close()
