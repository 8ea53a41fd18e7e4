import numpy as np, sys
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
for c in range(10):
    T = a["n_poses"][c]
    de = np.abs(a["est"][c, :T] - b["est"][c, :T]).max(axis=-1); dc = np.abs(a["cov"][c, :T] - b["cov"][c, :T]).max(axis=-1)
    print("clone", c, "T", T, "obs", int(a["obs"][c].sum()), "est diff", de.max(), "cov diff", dc.max(), "per pose cov", np.round(dc, 6), "meas/pose", np.diff(a["meas_ptr"][c][:T + 1]))
print("src obs", a["src_obs"].sum(axis=1))
