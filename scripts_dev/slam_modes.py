"""dev: which elimination mode k_slam takes per env-step (0 = rebuild from pose 0, 1 = light, 2 = rebuild from the checkpoint) over N ticks."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import bench
loop = bench.GpuLoop(0, 0, device_tick=False)
for _ in range(300):
    loop.tick()
rows = []
for _ in range(60):
    loop.tick()
    torch.cuda.synchronize()
    st = loop.env.eng.state
    clk = st["slam_clocks"].cpu().numpy(); act = st["active"].cpu().numpy().astype(bool)
    uc = st["update_count"].cpu().numpy()
    tot = (clk[:, 6] - clk[:, 0])
    for b in np.where(act)[0]:
        rows.append((int(clk[b, 10]), int(clk[b, 7]), int(uc[b]), float(tot[b]), float(clk[b, 2] - clk[b, 1])))
a = np.array(rows)
for m in (0, 1, 2):
    s = a[a[:, 0] == m]
    if len(s):
        print(f"mode {m}: {len(s)} env-steps, T mean {s[:, 1].mean():.1f} max {s[:, 1].max():.0f}; total cycles mean {s[:, 3].mean():.0f} max {s[:, 3].max():.0f}; B mean {s[:, 4].mean():.0f} max {s[:, 4].max():.0f}")
s = a[(a[:, 0] == 0) & (a[:, 1] > 12)]
print("mode-0 steps with T > 12:", len(s), "first rows (T, uc):", s[:12, 1:3].tolist())
