"""K1 standalone timing (no instrumentation), a few env variants."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
n, h, w = int(os.environ.get('NFR', 25)), 1024, 1280
pool = [torch.from_numpy(synth.bottle_batch(n, h, w, start_index=1000 * i)).cuda() for i in range(8)]
uni = [torch.full((n, h, w), 220, dtype=torch.uint8, device='cuda') for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
outs = [(torch.empty((n, h, w), dtype=torch.uint8, device='cuda'), torch.empty((n, h, w), dtype=torch.int32, device='cuda')) for _ in range(2)]
def run(tag, env, data):
    for k in ('HV_K1_DYNAMIC', 'HV_K1_SKIP_AUX', 'HV_NO_PDL', 'HV_K1_DEBUG_SKIP', 'HV_K1_L2PROM', 'HV_CCL_REPEAT'): os.environ.pop(k, None)
    os.environ.update(env)
    det = hc.Detector(0, profile=True); det.set_stream(st)
    def step(i): det.enqueue_device(data[i % 8].data_ptr(), n, h, w, 1, None, outs[i & 1][0].data_ptr(), outs[i & 1][1].data_ptr())
    for it in range(10): step(it)
    torch.cuda.synchronize(); det.profile()
    for it in range(50): step(it)
    torch.cuda.synchronize()
    pr = det.profile()
    k1 = pr['preprocess_mask']['ms'] / pr['preprocess_mask']['launches'] * 1e3
    ccl = pr['ccl_frame_fused']['ms'] / max(pr['ccl_frame_fused']['launches'], 1) * 1e3
    det.close()
    det = hc.Detector(0); det.set_stream(st)
    def step(i): det.enqueue_device(data[i % 8].data_ptr(), n, h, w, 1, None, outs[i & 1][0].data_ptr(), outs[i & 1][1].data_ptr())
    for it in range(20): step(it)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(200): step(it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    det.close()
    print('%-28s K1 %.1f us (%.2f of peak)  CCL %.1f us | step %.1f us  %.0f fps  frac %.3f' % (tag, k1, n*h*w*6/k1/1e3/6546.6, ccl, ms*1e3, n/ms*1e3, n*h*w*6/ms/1e6/6546.6))
for a in sys.argv[1:]:
    tag, _, rest = a.partition(':')
    env = dict(kv.split('=') for kv in rest.split(',') if kv)
    dk = env.pop('DATA', 'bottle')
    if dk == 'noise':
        rng = np.random.default_rng(0)
        data = [torch.from_numpy((128 + rng.integers(-8, 9, size=(n, h, w))).astype(np.uint8)).cuda() for _ in range(8)]
    else:
        data = uni if dk == 'uniform' else pool
    run(tag, env, data)
