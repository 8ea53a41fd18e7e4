"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this package.  The product package
``drl_graph_exploration_b200`` never does.  Parity status: the simulator / SLAM / virtual-map / graph / GCN path is PINNED
against the reference's own result files (all 200 episodes of data/test_result/{40,60,80,100}_DQN_GCN.csv: 6218 rows
reproduced, tests/golden/oracle_golden_scan.json; and, without their policies, 28 132 rows of the 1000 episodes of the other result
files -- A2C+GG-NN, Supervised+GCN, Nearest Frontier, Random, EM -- tests/golden/oracle_guided_scan.json); what those files do not exercise (roll-out rewards, GG-NN / g-U-Net layers)
is *parity unpinned* (see ``dge_oracle.hpp`` and DESIGN.md section 5).
"""
