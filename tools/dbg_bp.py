import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import housescan_b200 as hb
from housescan_b200 import synth
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
lib = ctx.lib
w, h = 640, 480
for nf in (1, 8, 16, 100, 1000):
    base, _ = synth.depth_stream(8, w, h)
    frames = torch.from_numpy(base.astype(np.int32)).to(dev).to(torch.int16).repeat((nf + 7) // 8, 1, 1)[:nf].contiguous()
    npx = nf * w * h
    bp_out = torch.zeros(npx * 3 + 16, dtype=torch.float32, device=dev)
    bp_cloud = ctx.wrap(bp_out.data_ptr(), npx, keepalive=bp_out)
    d_mask = torch.empty(npx, dtype=torch.uint8, device=dev)
    nv = C.c_int64()
    ctx._chk(lib.hs_backproject_ref_dev(ctx.h, C.c_void_p(frames.data_ptr()), w, h * nf, bp_cloud.h, C.c_void_p(d_mask.data_ptr()), C.byref(nv)))
    torch.cuda.synchronize()
    valid = frames.view(-1) != 0
    idx = torch.nonzero(valid).view(-1)
    d = (frames.view(-1)[idx].int() & 0xFFFF).float()
    exp = torch.stack([(idx % w).float() / 10.0, (idx // w).float() / 10.0, d / 20.0 - 30.0], dim=1)
    got = bp_out[: 3 * nv.value].view(-1, 3)
    cnt_ok = nv.value == int(valid.sum().item())
    mask_ok = bool(torch.equal(d_mask.bool(), valid))
    eq = (exp == got[: exp.shape[0]]) if cnt_ok else None
    print(nf, "count", cnt_ok, nv.value, int(valid.sum().item()), "mask", mask_ok, "points", None if eq is None else bool(eq.all().item()))
    if eq is not None and not bool(eq.all().item()):
        bad = torch.nonzero(~eq.all(dim=1)).view(-1)
        print("  first bad rows", bad[:5].tolist(), "of", bad.numel(), "exp", exp[bad[0]].tolist(), "got", got[bad[0]].tolist(), "pixel", int(idx[bad[0]]))
        cols = (~eq).sum(dim=0).tolist()
        print("  bad per column", cols)
