#!/usr/bin/env python3
"""Fixture for the feature-cache builder's pooling, produced by EXECUTING the reference's
``GlobalPooling2D`` (/root/reference/models/auxiliary/aux_models.py:54-64) on seeded inputs of the tap shapes the NTU
backbones produce (5-D visual maps, 4-D skeleton maps, 2-D vectors).

    python tests/golden/gen_golden_pooling.py        # writes tests/golden/pooling.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("MFAS_GOLDEN_OUT", HERE)          # tests regenerate into a scratch directory and compare
sys.path.insert(0, "/root/reference")
from models.auxiliary.aux_models import GlobalPooling2D  # noqa: E402

SHAPES = {"visual5d": (3, 32, 2, 7, 7), "ske4d": (4, 16, 5, 12), "vector": (5, 64), "odd": (2, 8, 147), "long": (2, 4, 4100)}


def main():
    g = torch.Generator().manual_seed(81)
    out = {}
    gp = GlobalPooling2D()
    for name, shape in SHAPES.items():
        x = torch.randn(*shape, generator=g).abs()
        out[name + "_x"] = x.numpy()
        out[name + "_y"] = gp(x).numpy()
    np.savez_compressed(os.path.join(OUT, "pooling.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
