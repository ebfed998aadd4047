"""Where does the CRNN logit error come from?  Feeds the GPU's rectified crops through the fp32 oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import receipts, weights
from oracle import ocr_ref
import torch.nn.functional as F
torch.set_num_threads(os.cpu_count())
sd = weights.crnn_calibrated(1, "CTC")
eng = bridge.Engine(act_dtype=0, head="CTC"); eng.load_state_dict(bridge.MODEL_CRNN, sd)
crops = receipts.crops(64, seed=21)
u8 = np.stack([ocr_ref.crop_to_tensor(g)[0] for g in crops])
out = eng.crnn_on_resized(u8)
rect_gpu = torch.from_numpy(eng.debug_read("rectified")).unsqueeze(1)
vis_gpu = torch.from_numpy(eng.debug_read("visual"))
with torch.no_grad():
    x = torch.cat([ocr_ref.crop_to_tensor(g)[1] for g in crops], 0)
    taps = {}
    ref = ocr_ref.crnn_forward(sd, x, "CTC", taps)
    def tail_from_rect(r):
        v = ocr_ref.resnet_features(sd, r)
        v = F.adaptive_avg_pool2d(v.permute(0, 3, 1, 2), (None, 1)).squeeze(3)
        s = ocr_ref._bilstm(sd, "SequenceModeling.0", v); s = ocr_ref._bilstm(sd, "SequenceModeling.1", s)
        return F.linear(s, sd["Prediction.weight"], sd["Prediction.bias"])
    def tail_from_vis(v):
        s = ocr_ref._bilstm(sd, "SequenceModeling.0", v); s = ocr_ref._bilstm(sd, "SequenceModeling.1", s)
        return F.linear(s, sd["Prediction.weight"], sd["Prediction.bias"])
    lg_rect = tail_from_rect(rect_gpu)
    lg_vis = tail_from_vis(vis_gpu)
gpu = torch.from_numpy(out["logits"])
def rep(name, a):
    d = (a - ref).abs()
    agree = (a.argmax(2) == ref.argmax(2)).float().mean().item()
    strs = np.mean([ocr_ref.ctc_decode(a[i].argmax(1)) == ocr_ref.ctc_decode(ref[i].argmax(1)) for i in range(len(crops))])
    print("%-44s max-abs %.4f mean-abs %.5f argmax agree %.4f strings %.3f" % (name, d.max().item(), d.mean().item(), agree, strs))
rep("GPU end to end", gpu)
rep("oracle fed with GPU rectified crops", lg_rect)
rep("oracle fed with GPU visual features", lg_vis)
print("rectified max-abs err", (rect_gpu - taps["rectified"]).abs().max().item(), "visual rel", ((vis_gpu - taps["visual"]).abs().max() / taps["visual"].abs().max()).item())
