"""Golden fixture for BrainTumorClassifier (main.py:301-328), produced by EXECUTING THE REFERENCE's own class (container-only):

    python tests/golden/make_golden_classifier.py

Weights/inputs come from the reference-independent recipes of oracle.unet3d_oracle; only the reference's outputs are stored
(tests/golden/classifier.npz).  Asserts that the oracle restatement agrees before writing."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_slice  # noqa: E402
from oracle import unet3d_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [("a", 2, (32, 32, 32), 21), ("b", 1, (16, 32, 48), 22), ("c", 1, (64, 64, 64), 23)]


def main():
    assert ref_slice.available(), "reference not mounted"
    ns = ref_slice.load()
    arrays = {}
    m = ns["BrainTumorClassifier"](4)
    arrays["keys"] = np.array(["%s %s" % (k, list(v.shape)) for k, v in m.state_dict().items()])
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == O.classifier_param_shapes(4)
    for name, n, (d, h, w), seed in CASES:
        sd = O.make_classifier_state_dict(4, seed=seed)
        x, _ = O.make_inputs(n, d, h, w, seed=seed)
        m.load_state_dict(sd)
        m.eval()
        with torch.no_grad():
            ref = m(x)
        got = O.classifier_forward(x, sd)
        assert float((ref - got).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max())), name
        arrays["logits_" + name] = ref.numpy()
    np.savez_compressed(os.path.join(OUT, "classifier.npz"), **arrays)
    print("wrote classifier.npz", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
