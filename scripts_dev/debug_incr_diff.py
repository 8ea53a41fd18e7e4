import numpy as np, sys
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
print("fro", a["fro"]); print("n_poses", a["n_poses"][:12], b["n_poses"][:12]); print("status", a["status"][:12], b["status"][:12]); print("uc", a["uc"][:12])
for k in ("raw", "est", "cov", "metrics", "est_l"):
    x, y = a[k], b[k]
    if k in ("est", "cov"):
        for c in range(12):
            T = a["n_poses"][c]
            d = np.abs(x[c, :T] - y[c, :T]).max(axis=-1)
            print(k, "clone", c, "T", T, "max diff", d.max(), "first bad pose", int(np.argmax(d > 1e-6)) if (d > 1e-6).any() else -1)
    else:
        print(k, np.abs(x - y).max())
print(a["raw"][0][:6], b["raw"][0][:6])
