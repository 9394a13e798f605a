"""World-size-2 gloo test of the multi-GPU host logic (bucket ownership + the one all_reduce), on CPU.
The per-rank "renderer" here is the oracle's sharded render (test infrastructure); the product path plugs
miro_gpu_render into the same render_sharded()."""
import os
import subprocess
import sys

import numpy as np

import helpers
from miro_b200 import distributed as md

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["MIRO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MIRO_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import helpers
from miro_b200 import distributed as md
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MIRO_PORT"], rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
fx = helpers.Fixture(helpers.fixture_path("c1_cornell")); sc = fx.scene()
p = sc.render_params(); p.width = p.height = 96
def fn(frame, si, sc_):
    p.shard_index, p.shard_count = si, sc_
    img, _ = helpers.oracle_render(sc, params=p)
    own = torch.from_numpy(md.bucket_owner(p.width, p.height, sc_) == si)
    frame[own] = torch.from_numpy(img)[own]
full = md.render_sharded(fn, p.width, p.height, rank, world)
sl = md.shard_rays(1001, rank, world)
cnt = torch.tensor([sl.stop - sl.start]); dist.all_reduce(cnt)
if rank == 0:
    np.save(os.environ["MIRO_OUT"], full.numpy()); assert int(cnt) == 1001
dist.destroy_process_group()
'''


def test_two_rank_sharded_render_equals_whole(tmp_path):
    out = tmp_path / "full.npy"
    env = dict(os.environ, MIRO_ROOT=helpers.ROOT, MIRO_PORT=str(29500 + os.getpid() % 2000), MIRO_OUT=str(out), WORLD_SIZE="2", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, "-c", WORKER], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell")); sc = fx.scene()
    p = sc.render_params(); p.width = p.height = 96
    whole, _ = helpers.oracle_render(sc, params=p)
    assert np.array_equal(np.load(out), whole)
    sc.close()


def test_bucket_ownership_partitions_the_image():
    for w, h, world in [(96, 96, 2), (1920, 1080, 8), (33, 65, 3), (16, 16, 4)]:
        own = md.bucket_owner(w, h, world)
        assert own.shape == (h, w) and own.min() == 0 and own.max() <= world - 1
        nbx = (w + 31) // 32
        assert own[0, 0] == 0 and (w <= 32 or own[0, 32] == 1 % world) and (h <= 32 or own[32, 0] == nbx % world)
    n = 10
    covered = np.zeros(n, int)
    for r in range(4):
        covered[md.shard_rays(n, r, 4)] += 1
    assert (covered == 1).all()
