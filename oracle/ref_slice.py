"""Load the UNMODIFIED reference classes by slicing their source (TEST INFRASTRUCTURE; container-only).

`import main` / `import training` cannot work: main.py:23-43 exits when flask/matplotlib/plotly are missing and
training.py:15 imports a non-existent name.  The model and loss classes themselves only need torch, so we exec exactly
their class definitions (located by `class` statements, not hard-coded line numbers) in a namespace holding torch, nn, F.
Nothing here is copied into the repo; /root/reference does not exist on the GPU box, so this module is used only by
tests/golden/make_golden.py (to produce fixtures) and by bench.py --impl reference / cpu_baseline when the reference
happens to be mounted.
"""
import importlib.util
import os
import re

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("B3D_REFERENCE_DIR", "/root/reference")


MATERIALISED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_classes.py")


def available():
    """True when the reference's own classes can be executed: /root/reference is mounted (build container) or oracle/_ref/ was
    materialised from it by oracle/make_ref.py (git-ignored, travels to the GPU box with the gpurun snapshot)."""
    return os.path.exists(os.path.join(REF, "main.py")) or os.path.exists(MATERIALISED)


def _load_materialised():
    spec = importlib.util.spec_from_file_location("_b3d_ref_classes", MATERIALISED)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return {k: getattr(mod, k) for k in dir(mod) if not k.startswith("__")}


def _class_block(src, name):
    lines = src.split("\n")
    start = next(i for i, l in enumerate(lines) if re.match(r"class %s\b" % name, l))
    end = start + 1
    while end < len(lines) and (lines[end].startswith((" ", "\t")) or lines[end].strip() == ""):
        end += 1
    return "\n".join(lines[start:end])


def load():
    """Returns a namespace dict with UNet3D, DoubleConv3D, AttentionGate3D (main.py), CombinedLoss3D, TverskyLoss3D,
    DeepSupervisionLoss3D (losses.py), CombinedLoss, DiceLoss, FocalLoss (training.py) and calculate_dice_score."""
    if not os.path.exists(os.path.join(REF, "main.py")):
        return _load_materialised()
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    main_src = open(os.path.join(REF, "main.py")).read()
    for cls in ("DoubleConv3D", "AttentionGate3D", "UNet3D", "BrainTumorClassifier"):
        exec(compile(_class_block(main_src, cls), "main.py:" + cls, "exec"), ns)
    tr_src = open(os.path.join(REF, "training.py")).read()
    for cls in ("DiceLoss", "FocalLoss", "CombinedLoss"):
        exec(compile(_class_block(tr_src, cls), "training.py:" + cls, "exec"), ns)
    # calculate_dice_score is a method of the trainer: slice the def and dedent it
    m = re.search(r"    def calculate_dice_score\(self, outputs, targets\):\n(?:(?:        .*|\s*)\n)+", tr_src)
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in m.group(0).split("\n"))
    exec(compile(body, "training.py:calculate_dice_score", "exec"), ns)
    try:   # input pipeline: three BraTSDataset methods whose bodies never touch `self`
        from scipy import ndimage
        ns["ndimage"] = ndimage
        for meth in ("_preprocess_image", "_preprocess_segmentation", "_apply_augmentations"):
            m = re.search(r"    def %s\(self.*?\):\n(?:(?:        .*|\s*)\n)+" % meth, tr_src)
            body = "\n".join(l[4:] if l.startswith("    ") else l for l in m.group(0).split("\n"))
            exec(compile(body, "training.py:" + meth, "exec"), ns)
    except ImportError:   # scipy missing: the model / loss classes above are still usable
        pass
    spec = importlib.util.spec_from_file_location("_ref_losses", os.path.join(REF, "losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for cls in ("CombinedLoss3D", "TverskyLoss3D", "DeepSupervisionLoss3D"):
        ns[cls] = getattr(mod, cls)
    return ns
