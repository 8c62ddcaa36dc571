"""Worker for tests/test_multi_gpu_cpu.py: one rank of a world_size-2 gloo job on CPU.  Each rank takes its shard of
a batch (contiguous slice), computes the per-frame results with the oracle standing in for the GPU (tests may use the
oracle), builds its line-statistics vector and all-reduces it exactly like bench.py does over NCCL."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "heimdall-vision_b200"))

import hv_dist  # noqa: E402
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def stats_vector(results):
    v = np.zeros(hv_dist.STATS_WORDS, np.int64)
    for r in results:
        v[0] += 1
        v[1] += 1 if r.defects else 0
        v[2] += len(r.defects)
        v[3] += r.ncomp
        v[4] += int(sum(d["size"] for d in r.defects))
        v[5] += int((r.mask == 255).sum())
        for d in r.defects:
            v[6 + min(int(d["size"]).bit_length() - 1, 15)] += 1
    return v


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, h, w = 6, 96, 128
    batch = synth.bottle_batch(n, h, w, start_index=900, contaminants=2)
    lo, hi = hv_dist.shard_range(n, rank, world)
    mine = [O.detect_contamination(batch[f][:, :, None]) for f in range(lo, hi)]
    vec = torch.from_numpy(stats_vector(mine))
    work = hv_dist.allreduce_line_stats(vec, async_op=True)
    if work is not None:
        work.wait()
    per_frame = {str(f): [[d["position"][0], d["position"][1], d["size"], d["confidence"]] for d in r.defects]
                 for f, r in zip(range(lo, hi), mine)}
    gathered = [None] * world
    dist.all_gather_object(gathered, per_frame)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        json.dump({"stats": vec.tolist(), "frames": merged, "world": world}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
