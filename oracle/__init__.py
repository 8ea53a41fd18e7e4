"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this package.  The product package
``drl_graph_exploration_b200`` never does.  Parity status: *parity unpinned* (see
``dge_oracle.hpp`` and DESIGN.md).
"""
