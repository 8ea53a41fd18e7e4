"""world_size-2 gloo test (CPU) of the N>1 path: env sharding and the single flat gradient all-reduce."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drl_graph_exploration_b200.dist import FlatGradBucket, env_seeds, shard_envs


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
    bucket = FlatGradBucket(model.parameters())
    lo, hi = shard_envs(10, rank, world)
    x = torch.arange(10 * 5, dtype=torch.float32).view(10, 5)[lo:hi] / 50.0
    bucket.zero_()
    model(x).sum().backward()
    local = bucket.flat.clone()
    bucket.all_reduce_mean()
    bucket.clamp_(0.5)
    q.put((rank, lo, hi, local.numpy().copy(), bucket.flat.numpy().copy()))   # by value: a tensor would travel as an fd the exiting worker may close first
    dist.destroy_process_group()


def test_shard_and_flat_allreduce():
    assert shard_envs(256 * 8, 3, 8) == (768, 1024) and shard_envs(10, 1, 4) == (3, 6) and shard_envs(10, 3, 4) == (9, 10)
    assert env_seeds(4, 6, episode=2).tolist() == [4 + 2 * (1 << 20), 5 + 2 * (1 << 20)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    (r0, lo0, hi0, l0, f0), (r1, lo1, hi1, l1, f1) = [(r, lo, hi, torch.from_numpy(l), torch.from_numpy(f)) for r, lo, hi, l, f in res]
    assert (lo0, hi0, lo1, hi1) == (0, 5, 5, 10)
    assert torch.equal(f0, f1)                                     # replicas stay identical
    assert torch.allclose(f0, ((l0 + l1) / 2).clamp(-0.5, 0.5))   # mean, THEN clamp


class _TinyQ(torch.nn.Module):
    """Stand-in with the Networks.GCN call signature (the CUDA GNN kernels need a GPU; the collective does not)."""

    def __init__(self):
        super().__init__()
        self.l1, self.l2 = torch.nn.Linear(5, 16), torch.nn.Linear(16, 1)

    def forward(self, data, prob, batch=None):
        return self.l2(torch.relu(self.l1(data.x)))


def _graphs(seed, n_graphs):
    from drl_graph_exploration_b200.data import Batch, Data
    g = torch.Generator().manual_seed(seed)
    items = [Data(torch.randn(int(n), 5, generator=g), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0)) for n in torch.randint(3, 9, (n_graphs,), generator=g)]
    return Batch.from_data_list(items)


def _train_worker(rank, world, port, q):
    from drl_graph_exploration_b200.policy import DeepQ
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _TinyQ()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    dq = DeepQ()
    for step in range(3):                       # DeepQ.train: forward, loss, backward, ONE all-reduce, clamp, Adam
        b = _graphs(100 * step + rank, 4)       # every rank trains on its own minibatch
        n = b.x.size(0)
        a = torch.zeros(n); a[::3] = 1.0
        y = torch.linspace(-1, 1, n) * a
        dq.train(b, a, y, torch.device("cpu"), model, opt)
    q.put((rank, torch.cat([p.detach().flatten() for p in model.parameters()]).numpy().copy()))
    dist.destroy_process_group()


def test_dqn_train_step_keeps_replicas_identical_and_equals_averaged_gradients():
    from drl_graph_exploration_b200.policy import DeepQ
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    res = [(r, torch.from_numpy(w)) for r, w in res]
    assert torch.equal(res[0][1], res[1][1])                       # replicas identical after 3 steps
    # single-process restatement: average the two ranks' gradients, clamp, Adam
    torch.manual_seed(0)
    model = _TinyQ()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    dq = DeepQ()
    for step in range(3):
        grads = []
        for rank in range(2):
            b = _graphs(100 * step + rank, 4)
            n = b.x.size(0)
            a = torch.zeros(n); a[::3] = 1.0
            y = torch.linspace(-1, 1, n) * a
            model.zero_grad()
            dq.cost(model(b, 0.5), y, a).backward()
            grads.append([p.grad.clone() for p in model.parameters()])
        for p, g0, g1 in zip(model.parameters(), *grads):
            p.grad = ((g0 + g1) / 2).clamp(-0.5, 0.5)
        opt.step()
    ref = torch.cat([p.detach().flatten() for p in model.parameters()])
    assert torch.allclose(res[0][1], ref, rtol=1e-5, atol=1e-7)


def _a2c_run(seed_env, seed_sampling, ticks):
    """VecA2CTrainer on the CPU stand-in engine of test_trainer_logic_cpu (same initial weights everywhere)."""
    import test_trainer_logic_cpu as tl
    from drl_graph_exploration_b200 import trainer as trainer_mod
    from drl_graph_exploration_b200.policy import A2C
    saved, trainer_mod._stream_ptr = trainer_mod._stream_ptr, lambda dev: None
    try:
        torch.manual_seed(0)
        actor, critic = tl._Actor(), tl._Critic()
        a2c = A2C(); a2c.nstep = 3
        tr = trainer_mod.VecA2CTrainer(tl._Env(8, seed_env), actor, critic, a2c=a2c, lr=1e-2, seed=seed_sampling)
        for _ in range(ticks):
            tr.tick()
    finally:
        trainer_mod._stream_ptr = saved
    return tr, torch.cat([p.detach().flatten() for p in tr.params]).numpy().copy()


def _a2c_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr, w_diff = _a2c_run(seed_env=10 + rank, seed_sampling=rank, ticks=60)     # every rank its own envs: segments close in different ticks
    diff = (tr.train_steps, tr.segments, w_diff)
    tr, w_same = _a2c_run(seed_env=5, seed_sampling=5, ticks=40)               # both ranks the same envs
    q.put((rank, diff, (tr.train_steps, tr.segments, w_same)))
    dist.destroy_process_group()


def test_a2c_trainer_segment_weighted_allreduce_keeps_replicas_identical():
    """trainer.VecA2CTrainer at world size 2: a rank joins the collective of every tick in which ANY rank closed a segment (so
    the gradient-step counts agree although the segment counts do not), replicas stay bit-identical, and two ranks that hold
    the same envs reproduce the single-process run (each contributes its loss weighted by its share of the segments)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_a2c_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    (_, (steps0, seg0, w0), (ssteps0, sseg0, sw0)), (_, (steps1, seg1, w1), (ssteps1, sseg1, sw1)) = res
    assert steps0 == steps1 and steps0 >= 5
    assert seg0 != seg1                                                        # the shards really differed
    assert torch.equal(torch.from_numpy(w0), torch.from_numpy(w1))
    tr, w_single = _a2c_run(seed_env=5, seed_sampling=5, ticks=40)              # no process group in this process: world = 1
    assert (ssteps0, sseg0) == (ssteps1, sseg1) == (tr.train_steps, tr.segments)
    assert torch.equal(torch.from_numpy(sw0), torch.from_numpy(sw1))
    assert torch.allclose(torch.from_numpy(sw0), torch.from_numpy(w_single), rtol=1e-5, atol=1e-7)


def _dqn_worker(rank, world, port, q):
    import test_trainer_logic_cpu as tl
    from drl_graph_exploration_b200 import trainer as trainer_mod
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    trainer_mod._stream_ptr = lambda dev: None
    torch.manual_seed(0)
    pol, tgt = tl._QNet(), tl._QNet()
    env = tl._Env(16 if rank == 0 else 3, seed=20 + rank)          # rank 1 collects its first minibatch much later than rank 0
    tr = trainer_mod.VecDQNTrainer(env, pol, tgt, replay_capacity=64, observe=0, lr=1e-2, seed=rank)
    tr.dqn.BATCH = 8
    first = None
    for i in range(90):
        tr.tick()
        if first is None and tr.train_steps > 0:
            first = (i, tr.replay.size)
    q.put((rank, tr.train_steps, first, torch.cat([p.detach().flatten() for p in pol.parameters()]).numpy().copy()))
    dist.destroy_process_group()


def test_dqn_trainer_ranks_agree_on_the_tick_learning_starts():
    """trainer.VecDQNTrainer at world size 2 with unequal shards: the start of learning (OBSERVE decisions passed, one minibatch
    of transitions in the rank's replay) is agreed on collectively, so no rank ever enters the gradient all-reduce alone: both
    ranks take their first gradient step in the same tick, the same number of steps overall, and the replicas stay identical."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dqn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    (_, steps0, first0, w0), (_, steps1, first1, w1) = res
    assert steps0 == steps1 and steps0 > 10
    assert first0[0] == first1[0]                     # same tick ...
    assert first0[1] > first1[1] >= 8                 # ... although rank 0 had its minibatch long before
    assert torch.equal(torch.from_numpy(w0), torch.from_numpy(w1))
