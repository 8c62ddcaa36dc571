"""N>1 host logic on CPU: world_size-2 gloo job == single process (sharding covers every frame exactly once, per-frame
results do not depend on the shard, the all-reduced line statistics equal the single-process totals)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_ranges_partition_exactly():
    import hv_dist
    for n in (0, 1, 5, 25, 26, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [hv_dist.shard_range(n, r, world) for r in range(world)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in spans) <= -(-n // world) if n else True
    with pytest.raises(ValueError):
        hv_dist.shard_range(10, 2, 2)
    # camera-stream affinity: 8 streams over 1/2/4/8 GPUs, every stream has exactly one owner
    for world in (1, 2, 4, 8):
        owners = [hv_dist.streams_of(r, world, 8) for r in range(world)]
        assert sorted(s for o in owners for s in o) == list(range(8))
        assert all(len(o) == 8 // world for o in owners)


def _run_world(world, tmp_path, port):
    out = tmp_path / f"out_{world}.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world),
               OMP_NUM_THREADS="1")
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py"), str(out)], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        try:
            stdout, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("gloo worker timed out")
        assert p.returncode == 0, stdout.decode()[-2000:]
    return json.load(open(out))


def test_world2_gloo_equals_single_process(tmp_path):
    one = _run_world(1, tmp_path, 29611)
    two = _run_world(2, tmp_path, 29612)
    assert two["world"] == 2
    assert one["frames"] == two["frames"]          # identical per-frame defect lists, whichever rank computed them
    assert one["stats"] == two["stats"]            # all-reduced totals == single-process totals
    import hv_dist
    d = hv_dist.stats_dict(np.array(two["stats"]))
    assert d["frames_inspected"] == 6 and d["total_defects"] == sum(len(v) for v in two["frames"].values())
    assert sum(d[f"area_hist_{i}"] for i in range(16)) == d["total_defects"]
