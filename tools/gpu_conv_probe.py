"""GPU probe for the tcgen05 conv kernel: each case runs in its own subprocess (a deadlocked kernel
only loses that case).  Usage: python tools/gpu_conv_probe.py [case_substring]"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    # name: (B,H,W,Cin,Cout,R,stride,dil,pad(t,b,l,r),bias,relu,res,res_shift)
    "1x1_small":        (1, 16, 24, 64, 64, 1, 1, 1, (0, 0, 0, 0), 0, 0, 0, 0),
    "1x1_bias_relu_res": (1, 16, 24, 256, 128, 1, 1, 1, (0, 0, 0, 0), 1, 1, 1, 0),
    "3x3_same":         (2, 20, 28, 64, 64, 3, 1, 1, (1, 1, 1, 1), 1, 1, 0, 0),
    "3x3_s2_pad10":     (1, 24, 32, 64, 64, 3, 2, 1, (1, 0, 1, 0), 1, 1, 0, 0),
    "3x3_dil2":         (1, 23, 40, 128, 128, 3, 1, 2, (2, 2, 2, 2), 1, 1, 0, 0),
    "3x3_s2_dil2":      (1, 46, 80, 64, 64, 3, 2, 2, (1, 0, 1, 0), 1, 1, 0, 0),
    "1x1_s2_crop":      (2, 24, 32, 128, 256, 1, 2, 1, (0, -1, 0, -1), 1, 0, 0, 0),
    "1x1_upsample_res": (1, 24, 32, 128, 256, 1, 1, 1, (0, 0, 0, 0), 1, 0, 1, 1),
    "3x3_bigK_N256":    (1, 23, 40, 512, 256, 3, 1, 1, (1, 1, 1, 1), 1, 1, 0, 0),
    "1x1_many_tiles":   (1, 200, 200, 64, 64, 1, 1, 1, (0, 0, 0, 0), 1, 0, 0, 0),
    "3x3_many_tiles":   (2, 96, 160, 64, 128, 3, 1, 1, (1, 1, 1, 1), 1, 1, 1, 0),
    "1x1_N1024":        (1, 46, 80, 256, 1024, 1, 1, 1, (0, 0, 0, 0), 1, 1, 1, 0),
}

CHILD = r'''
import sys, json, numpy as np, torch
import torch.nn.functional as F
from object_detection_tracking_b200 import engine
name, spec, impl, split, a_mode = json.loads(sys.argv[1])
B,H,W,Cin,Cout,R,stride,dil,pad,use_bias,relu,use_res,res_shift = spec
rng = np.random.default_rng(hash(name) % 1000)
x = rng.standard_normal((B,H,W,Cin)).astype(np.float32)
w = (rng.standard_normal((R,R,Cin,Cout)) / np.sqrt(R*R*Cin)).astype(np.float32)
bias = rng.standard_normal(Cout).astype(np.float32) if use_bias else None
pt,pb,pl,pr = pad
xt = torch.from_numpy(x).permute(0,3,1,2)
if pb < 0: xt = xt[:, :, :H+pb, :]
if pr < 0: xt = xt[:, :, :, :W+pr]
xt = F.pad(xt, (pl, max(pr,0), pt, max(pb,0)))
ref = F.conv2d(xt, torch.from_numpy(w).permute(3,2,0,1).contiguous(), None if bias is None else torch.from_numpy(bias), stride=stride, dilation=dil)
Ho, Wo = ref.shape[2:]
res = None
if use_res:
    rh, rw = ((Ho+1)//2, (Wo+1)//2) if res_shift else (Ho, Wo)
    res = rng.standard_normal((B,rh,rw,Cout)).astype(np.float32)
    rt = torch.from_numpy(res).permute(0,3,1,2)
    if res_shift: rt = rt.repeat_interleave(2,2).repeat_interleave(2,3)[:, :, :Ho, :Wo]
    ref = ref + rt
if relu: ref = torch.relu(ref)
ref = ref.permute(0,2,3,1).contiguous().numpy()
out = engine.op_conv2d(x, w, bias, res, stride=stride, dil=dil, pad=pad, relu=bool(relu), res_shift=res_shift, impl=impl, split=bool(split), a_mode=a_mode)
err = np.abs(out - ref)
scale = np.abs(ref).max()
info = dict(name=name, impl=impl, split=split, a_mode=a_mode, shape=list(out.shape), max_err=float(err.max()), ref_max=float(scale),
            rel=float(err.max()/scale), frac_bad=float((err > 1e-2*scale).mean()), nan=int(np.isnan(out).sum()))
if info["frac_bad"] > 0:
    bad = np.argwhere(err > 1e-2*scale)
    info["first_bad"] = bad[:6].tolist()
    o2 = out.reshape(-1, Cout); r2 = ref.reshape(-1, Cout); e2 = err.reshape(-1, Cout) > 1e-2*scale
    info["bad_rows_mod128"] = np.bincount(np.nonzero(e2.any(1))[0] % 128, minlength=128).tolist()[:16]
    info["bad_cols"] = np.nonzero(e2.any(0))[0][:16].tolist()
    info["out_zero_frac"] = float((out == 0).mean())
print("RESULT " + json.dumps(info))
'''


def main():
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    variants = [("tcgen05", 0, -1), ("tcgen05", 1, -1), ("simt", 1, -1)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    for name, spec in CASES.items():
        if filt and filt not in name:
            continue
        vs = list(variants)
        if spec[5] == 1 and spec[6] == 1 and spec[8] == (0, 0, 0, 0):
            vs.append(("tcgen05", 1, 1))     # plain 1x1 also through the im2col path
        for impl, split, a_mode in vs:
            arg = json.dumps([name, spec, impl, split, a_mode])
            try:
                r = subprocess.run([sys.executable, "-c", CHILD, arg], env=env, cwd=root, capture_output=True,
                                   text=True, timeout=120)
                lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                if lines:
                    print(lines[-1][7:], flush=True)
                else:
                    print(json.dumps(dict(name=name, impl=impl, split=split, a_mode=a_mode, rc=r.returncode,
                                          err=(r.stderr or r.stdout)[-600:])), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps(dict(name=name, impl=impl, split=split, a_mode=a_mode, timeout=True)), flush=True)


if __name__ == "__main__":
    main()
