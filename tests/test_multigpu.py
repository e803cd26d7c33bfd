"""Multi-GPU parity.

gpu part: needs >= 2 GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): tests/mgpu_worker.py under
torch.distributed.run, one rank per GPU, slab-partitioned Gibbs sweeps and MGMC V-cycles over NCCL, compared bitwise
with the one-GPU result.
cpu part: the host-side plumbing of the N > 1 path (slab partition = PETSc ownership ranges, id broadcast) under
world_size-2 gloo."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_slabs_tile_the_grid():
    import parmgmc_b200 as pmg
    for n, r, a in ((4097, 1, 2), (8193, 2, 2), (32769, 8, 2), (512, 8, 2), (67, 2, 1), (1025, 4, 2)):
        s = pmg.partition_slabs(n, r, a)
        assert s[0][0] == 0 and s[-1][1] == n and all(s[i][1] == s[i + 1][0] for i in range(r - 1))
        assert all(lo % a == 0 for lo, _ in s) and all(hi > lo for lo, hi in s)
    with pytest.raises(ValueError):
        pmg.partition_slabs(3, 4)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import parmgmc_b200 as pmg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
slabs = pmg.partition_slabs(8193, world)
mine = [slabs[rank]]
every = [None] * world
dist.all_gather_object(every, mine[0])
assert every == slabs, (every, slabs)
# the id rank 0 would get from pmg_comm_unique_id travels as a 128-byte object, exactly as bench.py broadcasts it
uid = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
assert uid[0] == bytes(range(128))
# weak-scaling geometry of bench.py: one (n-1)-row slab per rank, the last rank owns the closing row
n = 4097
ny = (n - 1) * world + 1
s = pmg.partition_slabs(ny, world)
assert s[rank][0] == rank * (n - 1) and (s[rank][1] - s[rank][0]) == (n - 1) + (1 if rank == world - 1 else 0)
dist.destroy_process_group()
print("gloo ok", rank)
"""


def test_host_side_plumbing_world_size_2_gloo(tmp_path):
    w = tmp_path / "gloo_worker.py"
    w.write_text(_GLOO_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29517", str(w), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("gloo ok") == 2


@pytest.mark.gpu
def test_multi_gpu_matches_single_gpu_bitwise():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 4 if ngpu >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29519",
                        os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-6000:]
    assert "FAIL" not in r.stdout and r.stdout.count("-> OK") >= 8
