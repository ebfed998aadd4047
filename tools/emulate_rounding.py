"""CPU study: which CRNN layers' 16-bit rounding (activations and weights) dominates the logit error?
Restates the oracle ResNet with a rounding hook per layer; layers listed in EXACT stay fp32 (= what a split-precision
hi+lo layer would deliver)."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200.synth import receipts, weights, specs
from oracle import ocr_ref
torch.set_num_threads(os.cpu_count())
trained = os.environ.get("TRAINED", "0") == "1"
sd = weights.crnn_calibrated(1, "CTC", trained=trained)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
crops = receipts.crops(N, seed=21)
x = torch.cat([ocr_ref.crop_to_tensor(g)[1] for g in crops], 0)
fe = specs.FE
h16 = lambda t: t.half().float()

def fold(p, bn):
    w = sd[p + ".weight"]
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + 1e-5)
    b = sd[bn + ".bias"] - sd[bn + ".running_mean"] * s
    return w * s.view(-1, 1, 1, 1), b

def run(exact_upto, round_act=True, round_w=True):
    """Layers with index < exact_upto are computed without rounding."""
    idx = [0]
    def cbr(p, bn, t, relu=True, res=None, **kw):
        w, b = fold(fe + p, fe + bn)
        i = idx[0]; idx[0] += 1
        ex = i < exact_upto
        if round_w and not ex: w = h16(w)
        y = F.conv2d(t, w, b, **kw)
        if res is not None: y = y + res
        if relu: y = F.relu(y)
        if round_act and not ex: y = h16(y)
        return y
    def layer(li, t):
        for i in range(specs.RESNET_BLOCKS[li]):
            p = "layer%d.%d." % (li, i)
            o = cbr(p + "conv1", p + "bn1", t, padding=1)
            r = t
            if (fe + p + "downsample.0.weight") in sd:
                r = cbr(p + "downsample.0", p + "downsample.1", t, relu=False)
            t = cbr(p + "conv2", p + "bn2", o, res=r, padding=1)
        return t
    with torch.no_grad():
        r = ocr_ref.tps_rectify(sd, x)
        h = cbr("conv0_1", "bn0_1", r, padding=1)
        h = cbr("conv0_2", "bn0_2", h, padding=1)
        h = F.max_pool2d(h, 2, 2)
        h = cbr("conv1", "bn1", layer(1, h), padding=1)
        h = F.max_pool2d(h, 2, 2)
        h = cbr("conv2", "bn2", layer(2, h), padding=1)
        h = F.max_pool2d(h, 2, (2, 1), (0, 1))
        h = cbr("conv3", "bn3", layer(3, h), padding=1)
        h = layer(4, h)
        h = cbr("conv4_1", "bn4_1", h, stride=(2, 1), padding=(0, 1))
        h = cbr("conv4_2", "bn4_2", h)
        nl = idx[0]
        v = h.permute(0, 3, 1, 2).squeeze(3)
        s = ocr_ref._bilstm(sd, "SequenceModeling.0", v)
        s = ocr_ref._bilstm(sd, "SequenceModeling.1", s)
        return F.linear(s, sd["Prediction.weight"], sd["Prediction.bias"]), v, nl

ref, vref, nl = run(10 ** 6)
print("layers:", nl, "logit std %.3f" % ref.std().item())
def rep(name, lg, v):
    d = (lg - ref).abs()
    agree = (lg.argmax(2) == ref.argmax(2)).float().mean().item()
    strs = np.mean([ocr_ref.ctc_decode(lg[i].argmax(1)) == ocr_ref.ctc_decode(ref[i].argmax(1)) for i in range(N)])
    print("%-40s logit max-abs %.4f mean %.5f | visual rel %.5f | argmax %.4f strings %.3f" %
          (name, d.max().item(), d.mean().item(), ((v - vref).abs().max() / vref.abs().max()).item(), agree, strs), flush=True)
for k in (0, 2, 6, 12, 18, 24, 29, 33):
    lg, v, _ = run(k)
    rep("exact layers < %d" % k, lg, v)
lg, v, _ = run(0, round_w=False); rep("activations only (weights exact)", lg, v)
lg, v, _ = run(0, round_act=False); rep("weights only (activations exact)", lg, v)
