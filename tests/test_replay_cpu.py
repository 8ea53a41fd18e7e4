"""Device replay (replay.GraphReplay) and the vectorised DQN targets (trainer.dqn_targets) on CPU tensors:
packing equals the reference's DataLoader collation (Batch.from_data_list), FIFO eviction equals the deque's popleft,
targets equal DeepQ.build_targets (the literal restatement of policy.py:153-178)."""
import numpy as np
import pytest
import torch

from drl_graph_exploration_b200.data import Batch, Data
from drl_graph_exploration_b200.policy import DeepQ
from drl_graph_exploration_b200.replay import GraphReplay
from drl_graph_exploration_b200.trainer import dqn_targets


def _graph(rng, n):
    e = int(rng.integers(0, 3 * n))
    ei = torch.tensor(rng.integers(0, n, (2, e)), dtype=torch.long)
    return Data(torch.tensor(rng.normal(size=(n, 5)), dtype=torch.float32), ei, torch.tensor(rng.uniform(0.1, 6, e), dtype=torch.float32))


def _round(rng, sizes):
    """A decision round as the engine emits it: packed batch + prefix sums + per-graph key/frontier sizes."""
    items = [_graph(rng, n) for n in sizes]
    b = Batch.from_data_list(items)
    # the engine's edges are grouped by graph; Batch.from_data_list keeps that order
    nptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    eptr = torch.tensor(np.concatenate([[0], np.cumsum([d.edge_attr.numel() for d in items])]), dtype=torch.int32)
    fro = torch.tensor([int(rng.integers(1, n + 1)) for n in sizes], dtype=torch.int32)
    key = torch.tensor(sizes, dtype=torch.int32) - fro
    return items, b, nptr, eptr, key, fro


def test_store_gather_equals_collation_and_ring_wraps():
    rng = np.random.default_rng(0)
    rp = GraphReplay(capacity=6, node_cap=40, edge_cap=120, device="cpu", slack=4)     # G = 10 slots
    stored = {}
    for rnd in range(5):
        sizes = [int(v) for v in rng.integers(1, 40, int(rng.integers(1, 5)))]
        items, b, nptr, eptr, key, fro = _round(rng, sizes)
        slots = rp.store_graphs(b.x, b.edge_index, b.edge_attr, b.batch, nptr, eptr, key, fro, len(sizes))
        for s, d, k, f in zip(slots.tolist(), items, key.tolist(), fro.tolist()):
            stored[s] = (d, k, f)                                  # later rounds overwrite re-used slots
        assert slots.tolist() == [(rp.allocated - len(sizes) + i) % rp.G for i in range(len(sizes))]
    pick = torch.tensor(list(stored.keys())[::-1][:7])
    got, n, off = rp.gather(pick)
    ref = Batch.from_data_list([stored[int(s)][0] for s in pick])
    assert torch.equal(got.x, ref.x) and torch.equal(got.edge_index, ref.edge_index)
    assert torch.equal(got.edge_attr, ref.edge_attr) and torch.equal(got.batch, ref.batch)
    assert n.tolist() == [stored[int(s)][0].x.size(0) for s in pick]
    assert rp.gk[pick].tolist() == [stored[int(s)][1] for s in pick] and rp.gf[pick].tolist() == [stored[int(s)][2] for s in pick]


def test_transition_ring_is_fifo_like_the_deque():
    rp = GraphReplay(capacity=5, node_cap=4, edge_cap=4, device="cpu", slack=3)
    z = torch.zeros(8, dtype=torch.int64)
    rp.gserial[:] = torch.arange(8)
    for lo in (0, 3, 6):                                           # 3 + 3 + 2 transitions into a ring of 5
        m = min(3, 8 - lo)
        ids = torch.arange(lo, lo + m)
        rp.append(z[:m], ids, ids.float(), z[:m], torch.zeros(m, dtype=torch.bool))
    assert rp.size == 5
    assert sorted(rp.t_a[:rp.capacity].tolist()) == [3, 4, 5, 6, 7]   # the three oldest were dropped (position `capacity` is append_masked's trash slot)
    s, a, r, s1, term = rp.sample(5, check=True)
    assert sorted(a.tolist()) == [3, 4, 5, 6, 7] and torch.equal(a.float(), r)
    rp.gserial[0] = 99                                             # the slot was re-allocated under a live transition
    rp.sample(5)                                                   # the test runs on every sample (device flag, no sync) ...
    with pytest.raises(RuntimeError, match="wrapped"):
        rp.assert_intact()                                         # ... and is read back on demand


def test_serial_of_s_t_travels_with_the_in_flight_transition():
    """a slot recycled BETWEEN storing s_t and closing its transition is only visible if the serial read at store time
    is the one recorded -- reading the ring at append time records the new occupant's serial and the check passes silently."""
    def ring():
        rp = GraphReplay(capacity=4, node_cap=4, edge_cap=4, device="cpu", slack=0)      # G = 4 slots
        x, ei, ea = torch.zeros(2, 5), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0)
        one = lambda: rp.store_graphs(x, ei, ea, torch.zeros(2, dtype=torch.long), torch.tensor([0, 2]), torch.tensor([0, 0]), torch.tensor([1]), torch.tensor([1]), 1)
        return rp, one
    for carry in (True, False):
        rp, one = ring()
        s = one(); ser = rp.gserial[s].clone()                     # s_t stored: slot 0, serial 0
        for _ in range(4):
            s1 = one()                                             # four more decisions elsewhere: slot 0 is recycled (serial 4)
        rp.append(s, torch.tensor([1]), torch.tensor([0.5]), s1, torch.tensor([False]), serial_s=ser if carry else None)
        rp.sample(1)
        if carry:
            with pytest.raises(RuntimeError, match="wrapped"):
                rp.assert_intact()
        else:
            rp.assert_intact()                                     # (the old behaviour: blind to it)


def test_vectorised_targets_equal_build_targets():
    rng = np.random.default_rng(1)
    dq = DeepQ()

    class Net(torch.nn.Module):            # target-net stand-in: Q = first feature
        def forward(self, data, prob, batch=None):
            return data.x[:, :1]

    k = 9
    mb, rp = [], GraphReplay(capacity=32, node_cap=24, edge_cap=80, device="cpu", slack=8)
    sizes = [int(v) for v in rng.integers(2, 24, 2 * k)]
    items, b, nptr, eptr, key, fro = _round(rng, sizes)
    slots = rp.store_graphs(b.x, b.edge_index, b.edge_attr, b.batch, nptr, eptr, key, fro, 2 * k)
    a_node, r, term = [], [], []
    for i in range(k):
        s, s1 = items[2 * i], items[2 * i + 1]
        an = int(key[2 * i]) + int(rng.integers(0, int(fro[2 * i])))
        onehot = np.zeros(s.x.size(0)); onehot[an] = 1
        rew, t = float(rng.uniform(-1, 1)), bool(rng.integers(0, 2))
        mb.append((s, onehot, rew, s1, t, int(fro[2 * i + 1])))
        a_node.append(an); r.append(rew); term.append(t)
    _, a_ref, y_ref = dq.build_targets(mb, torch.device("cpu"), Net())
    rp.append(slots[0::2], torch.tensor(a_node), torch.tensor(r), slots[1::2], torch.tensor(term))
    s, a, rr, s1, tt = rp.t_s[:k], rp.t_a[:k], rp.t_r[:k], rp.t_s1[:k], rp.t_term[:k]
    b_s, n_s, off_s = rp.gather(s)
    b_s1, n_s1, off_s1 = rp.gather(s1)
    q1 = Net()(b_s1, 0.0).view(-1)
    act, y = dqn_targets(q1, b_s1.batch, a, rr, tt, off_s, n_s1, off_s1, rp.gf[s1], b_s.x.size(0), dq.GAMMA)
    assert torch.equal(act, a_ref.float()) and torch.allclose(y, y_ref.float(), rtol=1e-6, atol=1e-7)


def test_vectorised_nstep_returns_equal_the_reference_loop():
    """trainer.nstep_returns (S segments at once) against A2C.nstep_batch (policy.py:361-393, one segment)."""
    from drl_graph_exploration_b200.policy import A2C
    from drl_graph_exploration_b200.trainer import nstep_returns
    rng = np.random.default_rng(5)
    S, n = 6, 7
    r = rng.uniform(-1, 1, (S, n)); term = rng.uniform(size=(S, n)) < 0.3; last = rng.normal(size=S); vals = rng.normal(size=(S, n))
    out = nstep_returns(torch.tensor(r), torch.tensor(term, dtype=torch.float64), torch.tensor(last), 0.99).numpy()
    g = Data(torch.zeros(2, 5), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0))
    for s in range(S):
        ac = A2C(); ac.nstep = n
        for i in range(n):
            ac.buffer.append((g, np.array([0.0, 1.0]), r[s, i], g, bool(term[s, i]), 1, vals[s, i]))
        _, a, mask, returns, adv = ac.nstep_batch(last[s])
        assert np.allclose(out[s], returns, rtol=1e-12, atol=1e-12)
        assert np.allclose(adv.reshape(n, 2)[:, 1], returns - vals[s])


def test_append_masked_places_the_selected_rows_like_append():
    """append_masked (one row per env + mask, cursors on the device, no host count) against append of the compacted rows, over
    several wraps of the ring; sampling stays below the live size while the ring fills."""
    g = torch.Generator().manual_seed(0)
    a, b = (GraphReplay(capacity=11, node_cap=4, edge_cap=4, device="cpu", slack=40) for _ in range(2))
    for rp in (a, b):
        rp.gserial[:] = torch.arange(rp.G)
    B = 7
    for it in range(12):
        mask = torch.rand(B, generator=g) < 0.45
        s, s1 = torch.randint(0, a.G, (B,), generator=g), torch.randint(0, a.G, (B,), generator=g)
        s = torch.where(mask, s, torch.full_like(s, -1))                       # rows that are not selected may hold -1 slots
        act, r, term = torch.randint(0, 9, (B,), generator=g), torch.rand(B, generator=g), torch.rand(B, generator=g) < 0.3
        ser, ser1 = a.gserial[s.clamp(min=0)], a.gserial[s1]
        cnt = a.append_masked(mask, s, act, r, s1, term, serial_s=ser, serial_s1=ser1)
        idx = mask.nonzero().view(-1)
        b.append(s[idx], act[idx], r[idx], s1[idx], term[idx], serial_s=ser[idx], serial_s1=ser1[idx])
        assert int(cnt) == idx.numel() and (a.size, a.head) == (b.size, b.head)
        C = a.capacity
        for name in ("t_s", "t_s1", "t_a", "t_r", "t_term", "t_serial"):
            assert torch.equal(getattr(a, name)[:C], getattr(b, name)[:C]), (it, name)
        if a.size >= 3:
            gen = torch.Generator().manual_seed(it)
            _, aa, rr, _, _ = a.sample(3, generator=gen)
            live = set(zip(a.t_a[:C].tolist()[:a.size] if a.size < C else a.t_a[:C].tolist(), [round(float(v), 6) for v in (a.t_r[:C][:a.size] if a.size < C else a.t_r[:C])]))
            assert len(set(zip(aa.tolist(), [round(float(v), 6) for v in rr]))) == 3          # distinct positions
            assert all((x, round(float(y), 6)) in live for x, y in zip(aa.tolist(), rr))     # ... among the live ones
    a.assert_intact()


def test_state_dict_round_trip_and_the_older_layout_without_the_trash_slot():
    """state_dict / load_state_dict carry the rings and the device cursors; a checkpoint whose transition arrays have `capacity` rows
    (written before append_masked's trash slot existed) still loads."""
    a = GraphReplay(capacity=6, node_cap=4, edge_cap=4, device="cpu", slack=4)
    a.gserial[:] = torch.arange(a.G)
    z = torch.zeros(4, dtype=torch.int64)
    a.append(z, torch.arange(4), torch.arange(4).float(), z + 1, torch.tensor([False, True, False, True]))
    sd = a.state_dict()
    b = GraphReplay(capacity=6, node_cap=4, edge_cap=4, device="cpu", slack=4)
    b.load_state_dict(sd)
    assert (b.size, b.head) == (4, 4) and torch.equal(b.t_a, a.t_a) and torch.equal(b.t_term, a.t_term)
    old = {k: (v[:6] if k.startswith("t_") else v) for k, v in sd.items()}
    c = GraphReplay(capacity=6, node_cap=4, edge_cap=4, device="cpu", slack=4)
    c.load_state_dict(old)
    assert (c.size, c.head) == (4, 4) and torch.equal(c.t_a[:6], a.t_a[:6]) and torch.equal(c.t_serial[:6], a.t_serial[:6])
    c.sample(3, check=True)
