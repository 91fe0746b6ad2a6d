"""N > 1 path on CPU: two gloo ranks, the image axis is sharded with no data-path collective; only the timing reduction
and the barrier use torch.distributed."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    import torch
    from attentionshift_b200 import parallel as P
    rank, world, local = P.init(backend='gloo')
    assert world == 2
    mine = list(P.shard_images(13, rank, world))
    # every rank works on its own images only; pretend the per-image result is a checksum
    local_sum = float(sum(i * i for i in mine))
    P.barrier()
    tmax = P.max_over_ranks([10.0 + rank, 5.0 - rank])
    tot = P.sum_over_ranks([local_sum, float(len(mine))])
    # packed logging reduction (base.py:185-216 semantics with one all-reduce): rank r holds losses scaled by (r + 1)
    w = torch.tensor(2.0, requires_grad=True)
    losses = dict(loss_cls=w * (rank + 1) * torch.ones(3), loss_bbox=[w * 0.5 * (rank + 1), w * 0.25 * (rank + 1) * torch.ones(2)],
                  acc=torch.tensor(10.0 * (rank + 1)))
    loss, logs = P.parse_losses(losses)
    loss.backward()
    if rank == 0:
        print(json.dumps(dict(tmax=tmax, tot=tot, mine=mine, logs=logs, local_loss=float(loss), grad=float(w.grad))))
    import torch.distributed as dist
    dist.destroy_process_group()
''')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gloo_ranks_shard_and_reduce(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res['tmax'] == [11.0, 5.0]                       # max over ranks of (10+rank, 5-rank)
    assert res['tot'] == [float(sum(i * i for i in range(13))), 13.0]   # the two shards partition the 13 images
    assert res['mine'] == list(range(7))
    # mean over the two ranks of (r+1) x {2.0, 1.0 + 0.5, 10}: keys in the reference's order, 'loss' last, 'acc' not in the loss
    assert list(res['logs']) == ['loss_cls', 'loss_bbox', 'acc', 'loss']
    assert res['logs'] == dict(loss_cls=3.0, loss_bbox=2.25, acc=15.0, loss=5.25)
    assert res['local_loss'] == 3.5 and res['grad'] == 1.75             # rank 0's own loss keeps its autograd graph


def test_shards_partition_the_batch():
    from attentionshift_b200.parallel import shard_images
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in shard_images(n, r, world)]
            assert seen == list(range(n))
            sizes = [len(shard_images(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


DDP_WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, 'tests'))
    import torch
    import torch.distributed as dist
    from attentionshift_b200 import blocks, training
    from attentionshift_b200 import parallel as P
    import test_training_wiring as W

    class Patch:                                   # pytest's monkeypatch interface, permanent
        def setattr(self, obj, name, val):
            setattr(obj, name, val)
    W._emulate_kernels(Patch())
    rank, world, local = P.init(backend='gloo')
    B, T, heads, C = 2, 21, 2, 128

    class Net(torch.nn.Module):                    # two blocks = two BlockFn nodes under one DDP wrapper
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.blocks = torch.nn.ModuleList([blocks.Block(C, heads, mlp_ratio=4.0, qkv_bias=True) for _ in range(2)])
        def forward(self, x):
            for blk in self.blocks:
                x, _ = training.block_forward(blk, x, B, T, heads, False)
            return x
    net = Net()
    ddp = torch.nn.parallel.DistributedDataParallel(net, broadcast_buffers=False, gradient_as_bucket_view=True)
    data = [torch.randn(B * T, C, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
    w = torch.randn(B * T, C, generator=torch.Generator().manual_seed(99))
    (ddp(data[rank]) * w).sum().backward()
    got = [p.grad.clone() for p in net.parameters()]
    # the same gradients without DDP: mean over both ranks' shards
    ref = None
    for r in range(world):
        net.zero_grad(set_to_none=True)
        (net(data[r]) * w).sum().backward()
        g = [p.grad.clone() for p in net.parameters()]
        ref = g if ref is None else [a + b for a, b in zip(ref, g)]
    err = max(float((a - b / world).abs().max() / (b / world).abs().max().clamp_min(1e-30)) for a, b in zip(got, ref))
    flat = torch.cat([g.reshape(-1) for g in got])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = bool(all(torch.equal(gathered[0], t) for t in gathered))
    if rank == 0:
        print(json.dumps(dict(err=err, same=same, n=len(got))))
    dist.destroy_process_group()
''')


def test_ddp_all_reduces_the_fused_block_gradients(tmp_path):
    """The DDP gradient all-reduce of the training path (mmdet/apis/train.py:96-100) on two gloo ranks: every block
    is one custom autograd node (``training.BlockFn``, device kernels replaced by their torch restatements from
    tests/test_training_wiring.py), DDP's hooks must still see every parameter gradient, average it over the ranks and leave
    both ranks with identical gradients."""
    script = tmp_path / 'ddp_worker.py'
    script.write_text(DDP_WORKER % (ROOT, ROOT))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                   OMP_NUM_THREADS='2')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res['n'] == 24 and res['same'] and res['err'] < 1e-5, res
