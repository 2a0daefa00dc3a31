"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (one flat-buffer all-reduce)
averages gradients across ranks, keeps parameter/gradient layouts (channels_last weights) and leaves frozen
parameters alone.  The synthetic-input generator of the package must equal the oracle's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from improving_segmentation_with_selfsupervised_depth_b200.parallel import GradSync
    torch.manual_seed(0)
    w1 = torch.nn.Parameter(torch.randn(8, 4, 3, 3).contiguous(memory_format=torch.channels_last))
    b1 = torch.nn.Parameter(torch.randn(8))
    frozen = torch.nn.Parameter(torch.randn(5), requires_grad=False)
    sync = GradSync([w1, b1, frozen])
    assert frozen.grad is None
    assert w1.grad.stride() == w1.stride() and w1.grad.shape == w1.shape
    # per-rank "backward": autograd accumulates into the pre-assigned views
    x = torch.full((2, 4, 5, 5), float(rank + 1))
    sync.zero()
    y = torch.nn.functional.conv2d(x, w1, b1).sum()
    y.backward()
    local = w1.grad.clone()
    flat = sync.reduce()
    assert flat.data_ptr() == sync.flat.data_ptr()
    # average over ranks: gradient is linear in the input scale (1 and 2) -> 1.5x rank-0's gradient
    base = local / float(rank + 1)
    torch.testing.assert_close(w1.grad, base * 1.5)
    torch.testing.assert_close(b1.grad, torch.full((8,), 2 * 3 * 3.0))
    out[rank] = float(w1.grad.abs().sum())
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert abs(out[0] - out[1]) < 1e-6 * abs(out[0])


def test_synthetic_inputs_match_oracle():
    import segsde_oracle as O
    from improving_segmentation_with_selfsupervised_depth_b200.synthetic import synthetic_inputs
    a = synthetic_inputs(2, 32, 64, seed=3, labels=True)
    b = O.synthetic_inputs(2, 32, 64, seed=3, labels=True)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k
