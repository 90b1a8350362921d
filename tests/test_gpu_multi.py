"""Hardware data-parallel gradient parity: the NCCL all-reduced gradient arenas of G shards against ONE
GPU running the whole global batch.  Needs >= 2 visible GPUs (skips on the single-GPU box); one process per GPU,
NCCL over NVLink, rendezvous on 127.0.0.1.  The gloo / oracle counterpart that runs without a GPU is
tests/test_distributed_cpu.py."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    from hetersumgraph_b200 import _lib
    from hetersumgraph_b200.dist import gradient_parity
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    _lib.require_device()
    rep = gradient_parity(rank, world, dev, lambda t: dist.all_reduce(t), n_global=64)
    if rank == 0:
        import json
        with open(out_path, "w") as f:
            json.dump(rep, f)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_all_reduced_shard_gradients_equal_single_gpu_gradients(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import json

    import torch.multiprocessing as mp
    out = str(tmp_path / "parity.json")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    rep = json.load(open(out))
    assert rep["ranks"] == world and rep["ok"], rep
