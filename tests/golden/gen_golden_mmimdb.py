#!/usr/bin/env python3
"""Fixture for the MM-IMDB multi-label head, produced by EXECUTING the reference's
``WeightedCrossEntropyWithLogits`` (/root/reference/models/auxiliary/aux_models.py:129-147; autograd supplies the
gradient) and the metric call of its dev loop (sklearn ``f1_score(average='samples')`` at sigmoid > 0.3,
models/search/train_searchable/mmimdb.py:84,101).  Run in the build container:

    python tests/golden/gen_golden_mmimdb.py        # writes tests/golden/mmimdb_head.npz
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("MFAS_GOLDEN_OUT", HERE)          # tests regenerate into a scratch directory and compare
sys.path.insert(0, "/root/reference")
from models.auxiliary.aux_models import WeightedCrossEntropyWithLogits      # noqa: E402
from sklearn.metrics import f1_score                                        # noqa: E402


def main():
    g = torch.Generator().manual_seed(23)
    out = {}
    for name, (B, C, scale) in {"a": (64, 23, 2.0), "b": (7, 23, 6.0), "c": (128, 23, 0.5)}.items():
        logits = (torch.randn(B, C, generator=g) * scale).requires_grad_(True)
        targets = (torch.rand(B, C, generator=g) < 0.15).float()
        if name == "b":
            targets[0] = 0                                  # a sample without labels: the 0/0 branch of the metric
        pos_weight = (torch.rand(C, generator=g) * 8 + 0.5).numpy().astype(np.float32)
        loss = WeightedCrossEntropyWithLogits(pos_weight)(logits, targets)
        loss.backward()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f1 = f1_score(targets.numpy(), (torch.sigmoid(logits.detach()) > 0.3).numpy(), average="samples")
        out.update({f"{name}_logits": logits.detach().numpy(), f"{name}_targets": targets.numpy(), f"{name}_pos_weight": pos_weight,
                    f"{name}_loss": np.float32(loss.item()), f"{name}_dlogits": logits.grad.numpy(), f"{name}_f1": np.float64(f1)})
    np.savez_compressed(os.path.join(OUT, "mmimdb_head.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if k.endswith(("loss", "f1"))})


if __name__ == "__main__":
    main()
