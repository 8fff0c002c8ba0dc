"""Global average pooling of a backbone tap: CPU restatement (TEST INFRASTRUCTURE ONLY; nothing under mfas_b200/ imports it).

Restates ``GlobalPooling2D.forward`` (/root/reference/models/auxiliary/aux_models.py:58-64): ``x.view(B, C, -1)``,
``torch.mean(x, 2)``, ``view(B, -1)`` -- the operation the reference applies to every selected tap
(/root/reference/models/search/ntu_searchable.py:224-225) and the feature-cache builder applies once per sample
(mfas_b200/cache_builder.py).

Parity status: PINNED -- tests/test_cache_builder.py::test_pooling_oracle_matches_reference_module compares it with the
reference class executed in the build container on committed shapes (the class is 6 lines of torch; the fixture is the
script-generated tests/golden/pooling.npz).
"""
import numpy as np


def global_pool(x):
    """x: [B, C, ...] -> [B, C] fp32 (accumulated in float64: the summation order of torch.mean is not part of the contract)"""
    x = np.asarray(x)
    B, C = x.shape[0], x.shape[1]
    return x.reshape(B, C, -1).astype(np.float64).mean(axis=2).astype(np.float32)
