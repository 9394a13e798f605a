"""Multi-GPU plumbing: one process per GPU (torch.distributed), scene replicated, image sharded by 32x32 bucket.

The data path needs no collective: rank r renders the buckets b with b % world == r (bucket order of the reference's
render loop, src/Scene.cpp:160-175) into a zero-initialised frame; pixels of other ranks stay zero.  The only exchange
is ONE all_reduce(SUM) of the w*h*3 float frame at the end (NCCL over NVLink on GPUs; gloo in the CPU tests).
Random numbers are keyed by pixel, so the combined frame equals the single-GPU frame.
"""
import numpy as np

BUCKET = 32


def bucket_owner(width, height, world):
    """(height, width) int array: the rank that renders each pixel."""
    ys, xs = np.mgrid[0:height, 0:width]
    nbx = (width + BUCKET - 1) // BUCKET
    return ((ys // BUCKET) * nbx + xs // BUCKET) % max(world, 1)


def render_sharded(render_fn, width, height, rank, world, group=None, device="cpu"):
    """render_fn(frame, shard_index, shard_count) fills this rank's pixels of `frame` (a zeroed (h, w, 3) float32
    torch tensor on `device`) and leaves the others untouched.  Returns the combined frame (on every rank)."""
    import torch
    import torch.distributed as dist
    frame = torch.zeros((height, width, 3), dtype=torch.float32, device=device)
    render_fn(frame, rank, world)
    if world > 1:
        dist.all_reduce(frame, op=dist.ReduceOp.SUM, group=group)
    return frame


def render_scene_distributed(scene, rank, world, group=None, mode="tiles"):
    """Scene::raytraceImage across the ranks of a process group: each rank's GPU renders its share (miro_gpu_render writing
    straight into the torch CUDA tensor), then one NCCL all_reduce(SUM).
    mode "tiles":   32x32 buckets round-robin (any configuration);
    mode "samples": every rank renders the whole frame with the paths p % world == rank of each camera sample, each weighted
                    1 / numPaths (path-traced configurations with min_subdivs == max_subdivs; better balanced when parts of
                    the image are empty).  The sum is the whole image in both modes (random numbers are keyed by pixel,
                    sample and path, not by rank)."""
    import torch
    p = scene.render_params()
    cam = scene.camera()

    def fn(frame, si, sc):
        if mode == "samples":
            p.path_shard_index, p.path_shard_count = si, sc
        else:
            p.shard_index, p.shard_count = si, sc
        scene.render_device(frame.data_ptr(), params=p, camera=cam)
        torch.cuda.synchronize()
    return render_sharded(fn, p.width, p.height, rank, world, group, device=torch.device("cuda", torch.cuda.current_device()))


def shard_rays(n, rank, world):
    """Contiguous slice of a ray batch traced by `rank` (bench.py: no collective, results stay on the producing GPU)."""
    per = (n + world - 1) // world
    return slice(min(rank * per, n), min((rank + 1) * per, n))
