"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE's own classes (container-only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only).  The fixtures pin oracle/unet3d_oracle.py (tests/test_oracle_golden.py) and are
the fixed-size parity targets of the CUDA path (tests/test_gpu_parity.py).  Weights/inputs come from the
reference-independent recipes in oracle.unet3d_oracle (make_state_dict / make_inputs), so only OUTPUTS are stored.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_slice  # noqa: E402
from oracle import unet3d_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(os.cpu_count())


def sample_idx(numel, k=256):
    g = np.random.RandomState(12345)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))


def summarize(t):
    t = t.detach().double().reshape(-1)
    idx = sample_idx(t.numel())
    return {"sum": float(t.sum()), "abs_sum": float(t.abs().sum()), "l2": float(t.norm()),
            "samples": [float(v) for v in t[torch.from_numpy(idx)]]}


def model_case(ns, name, features, n, size, seed, dropout, store_full):
    sd = O.make_state_dict(4, 4, features, seed=seed)
    x, y = O.make_inputs(n, size, size, size, seed=seed)
    model = ns["UNet3D"](4, 4, features=list(features), dropout_rate=dropout)
    model.load_state_dict(sd)
    rec = {"features": list(features), "n": n, "size": size, "seed": seed, "dropout": dropout}
    arrays = {}
    model.eval()
    with torch.no_grad():
        ev = model(x)
    rec["eval_logits"] = summarize(ev)
    if store_full:
        arrays["eval_logits"] = ev.numpy()
    rec["eval_argmax_counts"] = [int(v) for v in torch.bincount(ev.argmax(1).reshape(-1), minlength=4)]
    rec["eval_confusion"] = O.confusion_counts(ev, y).tolist()
    rec["eval_dice_score"] = ns["calculate_dice_score"](None, ev, y)
    model.train()
    if dropout > 0:
        torch.manual_seed(4242)
        masks = [F.dropout3d(torch.ones(n, f, 1, 1, 1), dropout, True).reshape(n, f) for f in features]
        arrays["dropout_masks"] = np.concatenate([m.numpy().reshape(-1) for m in masks])
        torch.manual_seed(4242)
    main, deep = model(x)
    crit = ns["DeepSupervisionLoss3D"]()
    loss = crit((main, deep), y)
    loss.backward()
    rec["train_main"] = summarize(main)
    rec["train_deep"] = [summarize(d) for d in deep]
    rec["ds_loss"] = float(loss)
    rec["combined3d_main"] = {k: float(v) for k, v in ns["CombinedLoss3D"]()(main.detach(), y)[1].items()}
    rec["trainer_combined_main"] = float(ns["CombinedLoss"]()(main.detach(), y))
    rec["bn_running_mean"] = [float(v) for v in model.final_conv[1].running_mean]
    rec["bn_running_var"] = [float(v) for v in model.final_conv[1].running_var]
    rec["grads"] = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            rec["grads"][k] = None
        else:
            g = p.grad.double()
            rec["grads"][k] = {"l2": float(g.norm()), "sum": float(g.sum())}
            if store_full and p.numel() <= 4096:
                arrays["grad." + k] = p.grad.numpy()
    if store_full:
        arrays["train_main"] = main.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    # oracle must agree with the reference before anything is written (pin)
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    om, od, bn = O.unet_forward(x, sdo, features, training=True,
                                dropout_masks=None if dropout == 0 else masks)
    ol = O.deep_supervision_loss(om, od, y)
    assert abs(float(ol) - float(loss)) < 1e-5 * max(1.0, abs(float(loss))), (float(ol), float(loss))
    assert (om - main).abs().max() < 1e-4, float((om - main).abs().max())
    print(name, "ds_loss", float(loss), "oracle", float(ol), "max|logit diff|", float((om - main).abs().max()))
    return rec


def loss_case(ns, seed, n, size):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(n, 4, size, size, size, generator=g) * 2.0).requires_grad_(True)
    deep = [(torch.randn(n, 4, size, size, size, generator=g) * 1.5).requires_grad_(True) for _ in range(4)]
    y = torch.randint(0, 4, (n, size, size, size), generator=g)
    rec = {"seed": seed, "n": n, "size": size}
    tot, parts = ns["CombinedLoss3D"]()(logits, y)
    tot.backward()
    rec["combined3d"] = parts
    arrays = {"combined3d_grad": logits.grad.numpy().copy()}
    logits.grad = None
    rec["tversky"] = float(ns["TverskyLoss3D"]()(logits, y))
    tl = ns["CombinedLoss"]()(logits, y)
    tl.backward()
    rec["trainer_combined"] = float(tl)
    arrays["trainer_grad"] = logits.grad.numpy().copy()
    logits.grad = None
    ds = ns["DeepSupervisionLoss3D"]()((logits, deep), y)
    ds.backward()
    rec["ds_loss"] = float(ds)
    arrays["ds_grad_main"] = logits.grad.numpy().copy()
    for i, d in enumerate(deep):
        arrays["ds_grad_deep%d" % i] = d.grad.numpy().copy() if d.grad is not None else np.zeros(1, np.float32)
    rec["ds_deep3_has_grad"] = deep[3].grad is not None
    rec["dice_score"] = ns["calculate_dice_score"](None, logits.detach(), y)
    rec["confusion"] = O.confusion_counts(logits.detach(), y).tolist()
    np.savez_compressed(os.path.join(OUT, "loss_seed%d.npz" % seed), **arrays)
    return rec


def block_cases(ns):
    rec, arrays = {}, {}
    # DoubleConv3D 16 -> 32 on 2x16x8^3
    sd_all = O.make_state_dict(16, 4, (32, 64, 128, 256, 512), seed=5)
    pre = "downs.0."
    dc = ns["DoubleConv3D"](16, 32)
    dc.load_state_dict({k[len(pre):]: v for k, v in sd_all.items() if k.startswith(pre)})
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 16, 8, 8, 8, generator=g)
    arrays["doubleconv_out"] = dc(x).detach().numpy()
    # AttentionGate3D(32, 32, 16)
    pre = "ups.13."
    ag = ns["AttentionGate3D"](32, 32, 16)
    ag.load_state_dict({k[len(pre):]: v for k, v in sd_all.items() if k.startswith(pre)})
    gg = torch.randn(2, 32, 8, 8, 8, generator=g)
    xx = torch.randn(2, 32, 8, 8, 8, generator=g)
    arrays["gate_out"] = ag(g=gg, x=xx).detach().numpy()
    rec["note"] = "inputs: Generator(9) randn x[2,16,8,8,8], then g[2,32,8,8,8], x[2,32,8,8,8]; weights make_state_dict(16,4,(32,64,64,64,64),seed=5) downs.0 / ups.13"
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **arrays)
    return rec


def init_case(ns, features):
    torch.manual_seed(0)
    m = ns["UNet3D"](4, 4, features=list(features))
    rec = {}
    for k, v in m.state_dict().items():
        rec[k] = hashlib.sha1(v.numpy().tobytes()).hexdigest()[:16]
    return rec


def main():
    assert ref_slice.available(), "reference not mounted"
    ns = ref_slice.load()
    torch.manual_seed(0)
    golden = {"torch": torch.__version__}
    for cfg, (cin, feats) in {"default4": (4, (32, 64, 128, 256, 512)), "default1": (1, (32, 64, 128, 256, 512)),
                              "light4": (4, (16, 32, 64, 128, 256))}.items():
        m = ns["UNet3D"](cin, 4, features=list(feats))
        golden["keys_" + cfg] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert golden["keys_" + cfg] == [[k, list(s)] for k, s in O.param_shapes(cin, 4, feats)]
        golden["nparams_" + cfg] = sum(p.numel() for p in m.parameters())
    golden["init_small"] = init_case(ns, (16, 32, 64, 128, 256))
    golden["model_small"] = model_case(ns, "model_small", (16, 32, 64, 128, 256), 1, 32, seed=1, dropout=0.0, store_full=True)
    golden["model_small_n2"] = model_case(ns, "model_small_n2", (16, 32, 64, 128, 256), 2, 32, seed=2, dropout=0.0, store_full=False)
    golden["model_small_dropout"] = model_case(ns, "model_small_dropout", (16, 32, 64, 128, 256), 2, 32, seed=3, dropout=0.2, store_full=False)
    golden["model_default"] = model_case(ns, "model_default", (32, 64, 128, 256, 512), 1, 32, seed=4, dropout=0.0, store_full=False)
    golden["loss"] = [loss_case(ns, 11, 2, 16), loss_case(ns, 12, 1, 8)]
    golden["blocks"] = block_cases(ns)
    with open(os.path.join(OUT, "golden.json"), "w") as fh:
        json.dump(golden, fh, indent=1)
    print("wrote", os.path.join(OUT, "golden.json"))


if __name__ == "__main__":
    main()
