"""ConvTranspose3d(k2,s2) fprop / dgrad at the level-0/1 boundary (2x64^3x64 <-> 2x128^3x32): time the planner's choice and
forced tile plans (env B3D_TD/TH/TW/KC/BN are read once per process, so every variant runs in a subprocess)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import b3d  # noqa
    from unet3d_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    lvl = int(os.environ.get("LVL", "0"))
    s, cin, cout = (64, 64, 32) if lvl == 0 else (32, 128, 64)
    x = torch.randn(2, s, s, s, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cin, cout, 2, 2, 2, device=dev) * 0.05
    b = torch.zeros(cout, device=dev)
    packs = ops.pack_weight_pair(w, True)
    dy = torch.randn(2, 2 * s, 2 * s, 2 * s, cout, device=dev).to(torch.bfloat16)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    which = os.environ.get("WHICH", "dgrad")
    if which == "dgrad":
        ms = t(lambda: ops.convT2_dgrad(dy, packs[ops.PACK_CONVT_DGRAD][0], packs[ops.PACK_CONVT_DGRAD][2], cin))
        gb = (dy.numel() + x.numel()) * 2 / 1e9
    else:
        ms = t(lambda: ops.convT2_fprop(x, packs[ops.PACK_CONVT_FPROP][0], b, cout))
        gb = (dy.numel() + x.numel()) * 2 / 1e9
    print("RESULT %s lvl%d %s: %.3f ms  %.0f GB/s" % (which, lvl, os.environ.get("TAG", "planner"), ms, gb / ms * 1e3), flush=True)
    sys.exit(0)
variants = [("planner", {})]
for kc in (16, 32, 64):
    for bn in (32, 64, 128, 256):
        variants.append(("KC%d BN%d" % (kc, bn), {"B3D_KC": str(kc), "B3D_BN": str(bn)}))
for th in (1, 2, 4, 8):
    variants.append(("TH%d" % th, {"B3D_TH": str(th)}))
for which in ("dgrad", "fprop"):
    for lvl in (0,):
        for tag, env in variants:
            e = dict(os.environ, WHICH=which, LVL=str(lvl), TAG=tag, B3D_VERBOSE="1", **env)
            r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=120)
            plan = [l for l in r.stderr.split("\n") if "[b3d] igemm" in l]
            res = [l for l in r.stdout.split("\n") if l.startswith("RESULT")]
            print((res[0] if res else "FAILED %s %s" % (which, tag)) + "   | " + (plan[-1][:170] if plan else r.stderr[-200:].replace("\n", " ")), flush=True)
