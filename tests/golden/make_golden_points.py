#!/usr/bin/env python
"""Golden vectors for the point-cloud tail of filter_depth, made by EXECUTING THE REFERENCE'S OWN LINES (build container
only: needs /root/reference).

    python tests/golden/make_golden_points.py     ->  tests/golden/points_kat.npz

eval.py cannot be imported (it parses the command line and imports plyfile at import time), and the statements live in
the middle of filter_depth, so the block eval.py:281-296 -- from `height, width = depth_est_averaged.shape[:2]` to
`vertex_colors.append(...)` -- is cut out of the source text, dedented and exec'd on seeded inputs.
"""
import os
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = open("/root/reference/eval.py").read().splitlines()


def reference_block():
    start = next(i for i, l in enumerate(SRC) if l.strip() == "height, width = depth_est_averaged.shape[:2]")
    end = next(i for i, l in enumerate(SRC) if l.strip().startswith("vertex_colors.append("))
    return textwrap.dedent("\n".join(SRC[start:end + 1]))


def main():
    rng = np.random.RandomState(3)
    h, w = 24, 40
    env = {"np": np, "vertexs": [], "vertex_colors": [],
           "depth_est_averaged": (500 + 300 * rng.rand(h, w)).astype(np.float64),
           "final_mask": rng.rand(h, w) > 0.4,
           "ref_img": rng.rand(h, w, 3).astype(np.float32),
           "ref_intrinsics": np.array([[361.5, 0, 20.0], [0, 360.4, 12.0], [0, 0, 1]], dtype=np.float32),
           "ref_extrinsics": np.array([[0.97, -0.05, 0.23, -120.0], [0.06, 0.99, -0.02, 15.0], [-0.23, 0.03, 0.97, 30.0],
                                       [0, 0, 0, 1]], dtype=np.float32)}
    block = reference_block()
    exec(block, env)
    g = {k: env[k] for k in ("depth_est_averaged", "final_mask", "ref_img", "ref_intrinsics", "ref_extrinsics")}
    g["vertices"] = env["vertexs"][0]
    g["colors"] = env["vertex_colors"][0]
    g["block_sha1"] = np.array(__import__("hashlib").sha1(block.encode()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "points_kat.npz"), **g)
    print("points_kat.npz", g["vertices"].shape, g["vertices"].dtype, g["colors"].dtype)


if __name__ == "__main__":
    main()
