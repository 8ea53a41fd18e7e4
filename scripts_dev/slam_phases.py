"""dev: in-situ phase timing of k_slam (SM clock at phase boundaries) after N ticks of the bench loop."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import bench

loop = bench.GpuLoop(0, 0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    loop.tick()
torch.cuda.synchronize()
clk = loop.env.eng.state["slam_clocks"].cpu().numpy()
act = loop.env.eng.state["active"].cpu().numpy().astype(bool)
T = clk[:, 7]
d = np.diff(clk[:, :7], axis=1).astype(np.float64)
names = ["A(linearise)", "B(fwd chain)", "S(schur gemm)", "C(inverse)", "D(bwd chain)", "E(marginals)"]
order = np.argsort(-T)
print("active envs", act.sum(), "T mean", T[act].mean(), "T max", T[act].max())
print("per-phase cycles: mean over active envs | slowest env | by T quantile")
for i, n in enumerate(names):
    print(f"{n:16s} mean {d[act, i].mean():10.0f}  max {d[act, i].max():10.0f}")
tot = d.sum(axis=1)
light = clk[:, 10] == 1
for name, sel in (("light (cached)", act & light), ("heavy (rebuild)", act & ~light)):
    if sel.any():
        print(f"-- {name}: {sel.sum()} envs, T mean {T[sel].mean():.1f} max {T[sel].max()}, n2 mean {clk[sel, 11].mean():.1f}; total mean {tot[sel].mean():.0f} max {tot[sel].max():.0f} cycles; "
              + " ".join(f"{n.split('(')[0]}={d[sel, i].mean():.0f}/{d[sel, i].max():.0f}" for i, n in enumerate(names)))
        top = np.argsort(-tot * sel)[:3]
        for b in top:
            print(f"     env {b} T {T[b]} n2 {clk[b, 11]} " + " ".join(f"{n.split('(')[0]}={d[b, i]:.0f}" for i, n in enumerate(names)), "total", tot[b])
print(f"{'total':16s} mean {tot[act].mean():10.0f}  max {tot[act].max():10.0f}   (1965 MHz: max = {tot[act].max() / 1965:.1f} us)")
for b in order[:4]:
    print(f"   inside phase B: pose recurrence (warp 0) {clk[b, 8] / T[b]:.0f} cycles/pose, border column 0 (incl. its waits on the recurrence) {clk[b, 9] / T[b]:.0f} cycles/pose")
    print("env", b, "T", T[b], " ".join(f"{n.split('(')[0]}={d[b, i]:.0f}" for i, n in enumerate(names)), "total", tot[b])
