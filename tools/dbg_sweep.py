import sys; sys.path.insert(0,'.')
import numpy as np
import cmdlmc_b200 as cm
from cmdlmc_b200 import runtime
from cmdlmc_b200.topology import DeviceTopology, build_with_retry
runtime.init(0)
rng = np.random.RandomState(2024)
for case in range(40):
    if case % 3 == 0:
        cell = rng.uniform(7.0, 16.0, size=3); hm = np.diag(cell)
    else:
        hm = np.diag(rng.uniform(8.0, 15.0, size=3))
        hm[1, 0], hm[2, 0], hm[2, 1] = rng.uniform(-0.45, 0.45, size=3) * np.array([hm[0, 0], hm[0, 0], hm[1, 1]])
        if case % 3 == 2:
            hm[0, 1], hm[0, 2], hm[1, 2] = rng.uniform(-1.5, 1.5, size=3)
        cell = hm.ravel()
    box = cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)
    heights = 1.0 / np.linalg.norm(np.linalg.inv(hm.T), axis=1)
    rc = float(rng.uniform(0.15, 0.5) * heights.min())
    if case % 4 == 1: rc = float(rng.uniform(0.55, 0.95) * heights.min())
    n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 97, 200, 333]))
    if case % 4 == 1: n = min(n, 97)
    p = rng.uniform(-0.6, 1.6, size=(n, 3)) @ hm
    if n > 3 and case % 5 == 0: p[1] = p[0]
    cutoff, buffer = 0.7 * rc, 0.3 * rc
    print("case", case, "n", n, "heights", heights.round(2), "rc", round(rc,3), "nc", np.floor(heights/rc), flush=True)
    t = build_with_retry(lambda cap: DeviceTopology(box, n, cutoff, buffer, 0, None, cap, path=1), p[None])
    print("   ok", int(t.frame_info()[0][0]), t.n_images, flush=True)
