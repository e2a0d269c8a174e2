"""CPU, world_size 2, gloo: the batch-sharded path.  Each rank computes the gradient of ITS shard's mean loss
(with the oracle as the compute, since there is no GPU here), FlatGradBucket averages them with one
all-reduce, and the result must equal the full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import convlstm_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from satflow_b200 import ConvLSTM
        from satflow_b200.distributed import FlatGradBucket, broadcast_parameters, shard_batch

        torch.manual_seed(100 + rank)  # different init per rank -> broadcast must fix it
        net = ConvLSTM(3, 4, 2)
        broadcast_parameters(net)
        names = [n for n, _ in net.named_parameters()]
        p = {n: t.detach().clone() for n, t in net.named_parameters()}
        g = torch.Generator().manual_seed(5)
        x = torch.randn(4, 2, 3, 6, 6, generator=g)
        tgt = torch.rand(4, 3, 2, 6, 6, generator=g)
        bucket = FlatGradBucket(net.parameters())
        bucket.zero_()
        xs, ts = shard_batch(x, rank, world), shard_batch(tgt, rank, world)
        y, sv = O.rollout_forward(xs, p, 3)
        _, dy = O.mse_loss_and_grad(y, ts)
        gl = O.rollout_backward(dy, sv, p)
        for n, prm in net.named_parameters():
            prm.grad.add_(gl[n])  # what autograd does with the C ABI's gradients on a GPU
        bucket.all_reduce_mean()
        if rank == 0:
            yf, svf = O.rollout_forward(x, p, 3)
            _, dyf = O.mse_loss_and_grad(yf, tgt)
            gf = O.rollout_backward(dyf, svf, p)
            errs = {n: O.rel_l2(dict(net.named_parameters())[n].grad, gf[n]) for n in names}
            q.put(("ok", errs, float(bucket.flat.numel())))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put(("err", repr(e), 0.0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gradients_equal_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, errs, n = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", errs
    assert n > 0
    for k, v in errs.items():
        assert v < 1e-5, (k, v)


def test_shard_batch_rejects_ragged():
    from satflow_b200.distributed import shard_batch

    with pytest.raises(ValueError):
        shard_batch(torch.zeros(5, 1), 0, 2)
    assert shard_batch(torch.arange(8).view(8, 1), 1, 4).flatten().tolist() == [2, 3]
