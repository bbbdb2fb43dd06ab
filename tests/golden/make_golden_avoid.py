#!/usr/bin/env python
"""Generates tests/golden/avoid/collision_dwa.npz from the UNMODIFIED reference sources compiled in
this container (oracle/_ref/libergodic_ref.so: collision.cpp, grid.cpp, dynamic_window.cpp,
numerics.hpp): Collision::collisionCheck, validate_control and both DynamicWindow::control
overloads on one seeded occupancy map.  /root/reference does not exist on the GPU box, so the
vectors are committed; re-run here to regenerate:

    python tests/golden/make_golden_avoid.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle.pyoracle import RefLib  # noqa: E402
from test_collision_cpu import random_map  # noqa: E402


def main():
    assert RefLib.available(), "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20261017)
    ys, xs, res, xmin, ymin = 140, 180, 0.05, -2.0, 1.0
    data = random_map(rng, ys, xs, p_occ=0.006)
    col = (0.2, 1.0, 0.05, 0.65)                                            # explore yaml-like radii
    dwa_cfg = (0.1, 2.0, 0.2, 2.5, 2.5, 1.0, 1.0, -1.0, 1.0, -1.0, 2.0, -2.0)  # explore_omni.yaml
    samples = (3, 8, 5)
    n = 400
    poses = np.column_stack([rng.uniform(xmin - 0.4, xmin + xs * res + 0.4, n),
                             rng.uniform(ymin - 0.4, ymin + ys * res + 0.4, n), rng.uniform(-np.pi, np.pi, n)])
    twists = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    twists[::6, 2] = 0.0
    inside = np.column_stack([rng.uniform(xmin + 0.3, xmin + xs * res - 0.3, n),
                              rng.uniform(ymin + 0.3, ymin + ys * res - 0.3, n), rng.uniform(-np.pi, np.pi, n)])
    vref = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    t = np.arange(50) * 0.1
    xt_ref = np.column_stack([xmin + 4.0 + 1.5 * np.cos(0.7 * t), ymin + 3.5 + 1.2 * np.sin(0.9 * t), 4.0 * np.sin(0.5 * t)])
    out = dict(data=data, res=res, xmin=xmin, ymin=ymin, col=np.array(col), dwa_cfg=np.array(dwa_cfg),
               samples=np.array(samples), poses=poses, twists=twists, inside=inside, vref=vref, xt_ref=xt_ref)
    out["hit"] = RefLib.collision_check(data, res, xmin, ymin, col, poses)
    out["valid_05"] = RefLib.validate_control(data, res, xmin, ymin, col, inside, twists, 0.1, 0.5)
    out["valid_20"] = RefLib.validate_control(data, res, xmin, ymin, col, inside, twists, 0.1, 2.0)
    out["dwa_found_twist"], out["dwa_u_twist"] = RefLib.dwa_control(data, res, xmin, ymin, col, dwa_cfg, samples, inside,
                                                                     twists, vref=vref)
    out["dwa_found_traj"], out["dwa_u_traj"] = RefLib.dwa_control(data, res, xmin, ymin, col, dwa_cfg, samples, inside,
                                                                   twists, xt_ref=xt_ref, dt_ref=0.1)
    path = os.path.join(HERE, "avoid", "collision_dwa.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB; hit {out['hit'].mean():.2f}, valid {out['valid_05'].mean():.2f} / "
          f"{out['valid_20'].mean():.2f}, dwa found {out['dwa_found_twist'].mean():.2f} / {out['dwa_found_traj'].mean():.2f}")


if __name__ == "__main__":
    main()
