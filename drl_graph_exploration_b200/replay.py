"""Device-resident replay memory for the DQN path.

The reference keeps a ``deque`` of ``(s_t, a_t, r_t, s_t1, terminal, fro_size1)`` tuples whose graphs are
``torch_geometric.data.Data`` objects on the host (policy.py:131-133) and re-collates 2 x 64 of them through a
``DataLoader`` for every gradient step (policy.py:141-151).  Here the graphs never leave HBM:

* **graph ring**: fixed-stride slots ``x [G, Ncap, 5] f32``, ``ei [G, 2, Ecap] i32`` (node ids local to the graph),
  ``ea [G, Ecap] f32`` plus per-slot ``(n_nodes, n_edges, key_size, fro_size)``.  A decision round of the vectorised
  trainer stores its whole graph batch with three indexed writes (no per-graph host work).  ``s_t1`` of one transition
  is ``s_t`` of the env's next one, so every graph is stored once.
* **transition ring**: ``(slot_s, action_node, reward, slot_s1, terminal)`` as flat device tensors; FIFO like the
  reference's ``popleft`` (policy.py:132-133).
* **minibatch**: ``gather(slots)`` packs k stored graphs into one PyG-layout batch (``x``, ``edge_index`` with node
  offsets, ``edge_attr``, ``batch``) with index arithmetic on the device -- one host sync for the two totals.

A slot is re-used after ``G`` allocations; ``G = capacity + slack`` with ``slack`` >= the number of graphs that can be
allocated between a transition's ``s_t`` and its completion.  Every transition carries the allocation serials of its two
graphs AS OF THE TIME THEY WERE STORED (the caller keeps the serial of ``s_t`` next to its slot while the transition is in
flight); ``sample`` compares them with the ring on every call and accumulates the verdict in a device flag
(``wrapped``, no host sync) that ``assert_intact`` reads.
The ring's cursors (``size``, ``head``) live on the DEVICE: ``append_masked`` takes one row per env plus a mask and places the
selected rows behind the head with a prefix sum, ``sample`` draws distinct positions below the live size with a masked top-k --
neither needs the host to know how many transitions a tick closed, so a training tick appends and samples without a host
synchronisation.  ``size`` / ``head`` read as plain ints on demand (a sync; logging, checkpoints, tests).
Everything is plain tensor plumbing, so the CPU tests exercise the same code.
"""
from __future__ import annotations

import torch

from .data import Batch


class GraphReplay:
    def __init__(self, capacity: int, node_cap: int, edge_cap: int, device, slack: int = 0):
        self.capacity, self.node_cap, self.edge_cap = int(capacity), int(node_cap), int(edge_cap)
        self.G = G = self.capacity + int(slack)
        dev = self.device = torch.device(device)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        self.x, self.ei, self.ea = z((G, node_cap, 5), torch.float32), z((G, 2, edge_cap), torch.int32), z((G, edge_cap), torch.float32)
        self.gn, self.ge, self.gk, self.gf = (z((G,), torch.int64) for _ in range(4))
        self.gserial = torch.full((G,), -1, dtype=torch.int64, device=dev)
        C = self.capacity
        # (position C of every transition array is a trash slot: append_masked sends the rows it does not select there)
        self.t_s, self.t_s1, self.t_a = z((C + 1,), torch.int64), z((C + 1,), torch.int64), z((C + 1,), torch.int64)
        self.t_r, self.t_term = z((C + 1,), torch.float32), z((C + 1,), torch.bool)
        self.t_serial = z((C + 1, 2), torch.int64)  # allocation serials of (s, s1) at append time: detects slot re-use
        self.wrapped = torch.zeros((), dtype=torch.bool, device=dev)   # a sampled transition referred to an overwritten graph
        self._size = torch.zeros((), dtype=torch.int64, device=dev)    # live transitions
        self._head = torch.zeros((), dtype=torch.int64, device=dev)    # next transition position
        self._size_lb = 0      # host-side lower bound of size (size never shrinks): lets sample() skip the read once it is >= k
        self.allocated = 0     # graphs allocated so far (host); slot of allocation i is i % G

    @property
    def size(self) -> int:
        """live transitions (reads the device cursor: a host sync)"""
        self._size_lb = int(self._size)
        return self._size_lb

    @size.setter
    def size(self, v: int):
        self._size.fill_(int(v)); self._size_lb = int(v)

    @property
    def head(self) -> int:
        return int(self._head)

    @head.setter
    def head(self, v: int):
        self._head.fill_(int(v))

    # ------------------------------------------------------------------ graphs ---
    def store_graphs(self, x, edge_index, edge_attr, batch, node_ptr, edge_ptr, key_size, fro_size, n_graphs: int) -> torch.Tensor:
        """Stores a packed batch of ``n_graphs`` graphs (PyG DataLoader layout; ``node_ptr`` / ``edge_ptr`` [>= n_graphs+1]
        exclusive prefix sums; edges of a graph contiguous and in graph order) and returns their slots [n_graphs] i64."""
        dev, G = self.device, self.G
        ng = int(n_graphs)
        serial = torch.arange(self.allocated, self.allocated + ng, device=dev)
        slots = serial % G
        self.allocated += ng
        nptr, eptr = node_ptr[:ng + 1].long(), edge_ptr[:ng + 1].long()
        n_nodes, n_edges = x.size(0), edge_attr.size(0)
        gid_n = batch[:n_nodes].long()
        loc_n = torch.arange(n_nodes, device=dev) - nptr[gid_n]
        self.x[slots[gid_n], loc_n] = x
        gid_e = gid_n[edge_index[0]]                       # graph of an edge = graph of its source node
        loc_e = torch.arange(n_edges, device=dev) - eptr[gid_e]
        se = slots[gid_e]
        self.ei[se, 0, loc_e] = (edge_index[0] - nptr[gid_e]).int()
        self.ei[se, 1, loc_e] = (edge_index[1] - nptr[gid_e]).int()
        self.ea[se, loc_e] = edge_attr
        self.gn[slots] = nptr[1:] - nptr[:-1]
        self.ge[slots] = eptr[1:] - eptr[:-1]
        self.gk[slots] = key_size[:ng].long()
        self.gf[slots] = fro_size[:ng].long()
        self.gserial[slots] = serial
        return slots

    def gather(self, slots: torch.Tensor):
        """Packs the graphs in ``slots`` [k] into one batch.  Returns (Batch, n_nodes [k], node_offset [k])."""
        dev, k = self.device, slots.numel()
        n, e = self.gn[slots], self.ge[slots]
        noff, eoff = torch.cumsum(n, 0) - n, torch.cumsum(e, 0) - e
        N, E = (int(v) for v in torch.stack([n.sum(), e.sum()]).tolist())       # the one host sync
        ar = torch.arange(k, device=dev)
        gid_n = torch.repeat_interleave(ar, n, output_size=N)
        loc_n = torch.arange(N, device=dev) - noff[gid_n]
        x = self.x[slots[gid_n], loc_n]
        gid_e = torch.repeat_interleave(ar, e, output_size=E)
        loc_e = torch.arange(E, device=dev) - eoff[gid_e]
        se = slots[gid_e]
        ei = torch.stack([self.ei[se, 0, loc_e], self.ei[se, 1, loc_e]]).long() + noff[gid_e]
        b = Batch(x, ei, self.ea[se, loc_e], gid_n)
        b.num_graphs = k
        return b, n, noff

    # ------------------------------------------------------------- transitions ---
    def append(self, slot_s, action_node, reward, slot_s1, terminal, serial_s=None, serial_s1=None):
        """Appends m transitions (device tensors [m]); the oldest are overwritten beyond ``capacity``.  ``serial_s`` /
        ``serial_s1``: allocation serials of the two graphs read when they were stored (``gserial[slot]`` right after
        ``store_graphs``); without them the ring's current serials are recorded, which cannot see a slot that was recycled
        while the transition was in flight."""
        m = int(slot_s.numel())
        if m == 0:
            return
        C, dev = self.capacity, self.device
        pos = (self._head + torch.arange(m, device=dev)) % C
        self._write(pos, slot_s, action_node, reward, slot_s1, terminal, serial_s, serial_s1)
        self._head = (self._head + m) % C
        self._size = torch.clamp(self._size + m, max=C)
        self._size_lb = min(C, self._size_lb + m)

    def append_masked(self, mask, slot_s, action_node, reward, slot_s1, terminal, serial_s=None, serial_s1=None):
        """``append`` of the rows selected by ``mask`` (all arguments [B] device tensors, one row per env), in row order, WITHOUT the
        host learning how many there are: row b goes to ``head + (number of selected rows before b)``, unselected rows to the trash
        slot.  B must not exceed the capacity.  Returns the number of appended rows as a device scalar."""
        C = self.capacity
        mask = mask.bool()
        assert mask.numel() <= C, "append_masked: more rows than the ring holds"
        rank = torch.cumsum(mask.long(), 0) - 1
        pos = torch.where(mask, (self._head + rank) % C, torch.full_like(rank, C))
        self._write(pos, slot_s.clamp(min=0), action_node, reward, slot_s1.clamp(min=0), terminal, serial_s, serial_s1)
        cnt = mask.long().sum()
        self._head = (self._head + cnt) % C
        self._size = torch.clamp(self._size + cnt, max=C)
        return cnt

    def _write(self, pos, slot_s, action_node, reward, slot_s1, terminal, serial_s, serial_s1):
        self.t_s[pos], self.t_s1[pos], self.t_a[pos] = slot_s.long(), slot_s1.long(), action_node.long()
        self.t_r[pos], self.t_term[pos] = reward.float(), terminal.bool()
        self.t_serial[pos, 0] = self.gserial[slot_s.long()] if serial_s is None else serial_s.long()
        self.t_serial[pos, 1] = self.gserial[slot_s1.long()] if serial_s1 is None else serial_s1.long()

    def sample(self, k: int, generator=None, check: bool = False):
        """k distinct transitions (random.sample, policy.py:141) -> (slot_s, action_node, reward, slot_s1, terminal)."""
        if self._size_lb < k:                    # (until the host has once seen size >= k; size never shrinks)
            assert self.size >= k, "replay holds fewer transitions than the minibatch"
        # k distinct positions below the live size, which stays on the device: random keys, dead positions pushed out, k smallest
        key = torch.rand(self.capacity, device=self.device, generator=generator)
        key = torch.where(torch.arange(self.capacity, device=self.device) < self._size, key, torch.full_like(key, 2.0))
        idx = torch.topk(key, k, largest=False).indices
        s, s1 = self.t_s[idx], self.t_s1[idx]
        # a stored graph was overwritten while a live transition still refers to it (slack too small): always tested, on the device
        ok = (self.gserial[s] == self.t_serial[idx, 0]) & (self.gserial[s1] == self.t_serial[idx, 1])
        self.wrapped |= ~ok.all()
        if check:
            self.assert_intact()
        return s, self.t_a[idx], self.t_r[idx], s1, self.t_term[idx]

    def assert_intact(self):
        """Raises if any minibatch sampled so far contained a transition whose graph slot had been recycled (one host sync)."""
        if bool(self.wrapped):
            raise RuntimeError("graph ring wrapped over a live transition: the replay trained on a wrong graph -- increase `slack`")

    _TENSORS = ("x", "ei", "ea", "gn", "ge", "gk", "gf", "gserial", "t_s", "t_s1", "t_a", "t_r", "t_term", "t_serial")

    def state_dict(self) -> dict:
        """Everything needed to resume: the two rings and their cursors (the reference pickles its deque of Data objects,
        train.py:33-35 / run_training.py:63-64)."""
        d = {k: getattr(self, k).detach().cpu() for k in self._TENSORS}
        d.update(size=self.size, head=self.head, allocated=self.allocated, capacity=self.capacity, G=self.G,
                 node_cap=self.node_cap, edge_cap=self.edge_cap)
        return d

    def load_state_dict(self, d: dict):
        if (d["capacity"], d["G"], d["node_cap"], d["edge_cap"]) != (self.capacity, self.G, self.node_cap, self.edge_cap):
            raise ValueError("replay checkpoint was written with other capacities")
        for k in self._TENSORS:
            t, src = getattr(self, k), d[k]
            if k.startswith("t_") and src.shape[0] == self.capacity:      # a checkpoint written before the trash slot existed
                t[:self.capacity].copy_(src)
            else:
                t.copy_(src)
        self.size, self.head, self.allocated = int(d["size"]), int(d["head"]), int(d["allocated"])      # (the setters fill the device cursors)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x, self.ei, self.ea))
