"""N > 1 path on CPU (gloo, world_size 2): chain sharding, Philox stream offsets and the max-over-ranks timing
reduction used by bench.py -- no data-path collective exists, so this covers the host-side rank logic."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gpfanova_b200 import synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 64
    C = total // world
    first = rank * C
    start = synthetic.chain_starts(C, 4, first=first)
    # max-over-ranks timing exactly as bench.py does it
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros((C, 9), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.as_tensor(start))
    if rank == 0:
        np.save(out, np.concatenate([g.numpy() for g in gathered]))
        assert t.item() == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()


def test_chain_sharding_is_independent_of_world_size(tmp_path):
    import torch.multiprocessing as mp
    from gpfanova_b200 import synthetic
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "starts.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sharded = np.load(out)
    single = synthetic.chain_starts(64, 4, first=0)
    assert np.array_equal(sharded, single)          # rank r, local chain c == global chain r*C + c


def test_bench_reference_arm_nonzero_ranks_exit_quietly():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_chain_blocks_for_devices():
    """engine.split_chains: the partition EngineGroup (devices=[...]) and a torchrun launch both use"""
    from gpfanova_b200.engine import split_chains
    assert split_chains(4096, 8) == ([512] * 8, [512 * i for i in range(8)])
    counts, offsets = split_chains(10, 4)
    assert counts == [3, 3, 2, 2] and offsets == [0, 3, 6, 8]
    assert split_chains(5, 5) == ([1] * 5, list(range(5)))
    import pytest
    with pytest.raises(ValueError):
        split_chains(3, 4)
