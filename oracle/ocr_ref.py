"""TEST INFRASTRUCTURE - CPU oracle: the reference's detect-then-recognize path restated function by function.

The reference is Python that delegates its arithmetic to torch / cv2 / PIL (none of which live under
/root/reference), so this restatement calls the same third-party routines on the CPU in fp32 and re-expresses the
reference's own glue.  It is written functionally over a plain state dict (no nn.Module mirror of the reference
classes) and is pinned against the live reference by tests/golden/ (see oracle/make_golden.py).

Every function cites the reference lines it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import this module.
"""
import math
from functools import cmp_to_key

import cv2
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

from . import specs

_EPS = 1e-5
BN_CALIBRATE = False  # set by oracle/weights.build_calibrations only: record batch statistics into the state dict


def _bn(sd, p, x):
    if BN_CALIBRATE:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            True, 1.0, _EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, _EPS)


def _conv(sd, p, x, **kw):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), **kw)


# ------------------------------------------------------------------------------------------------ CRAFT
def craft_preproc(image, canvas_size=1280, mag_ratio=1.5):
    """resizeAspectRatio (tools/imgproc.py:38-65) + normalizeMeanVariance (:19-25) + CRAFT.preproc (net.py:71-80).
    image: uint8 BGR [H,W,3].  Returns (x [1,3,H32,W32] fp32, ratio_w, ratio_h)."""
    h, w, ch = image.shape
    target = min(mag_ratio * max(h, w), canvas_size)
    ratio = target / max(h, w)
    th, tw = int(h * ratio), int(w * ratio)
    proc = cv2.resize(image, (tw, th), interpolation=cv2.INTER_LINEAR)
    h32 = th if th % 32 == 0 else th + (32 - th % 32)
    w32 = tw if tw % 32 == 0 else tw + (32 - tw % 32)
    canvas = np.zeros((h32, w32, ch), np.float32)
    canvas[:th, :tw, :] = proc
    # ImageNet RGB constants applied to BGR channel order exactly as the reference does
    mean = np.array([0.485 * 255.0, 0.456 * 255.0, 0.406 * 255.0], np.float32)
    std = np.array([0.229 * 255.0, 0.224 * 255.0, 0.225 * 255.0], np.float32)
    canvas -= mean
    canvas /= std
    x = torch.from_numpy(canvas).permute(2, 0, 1).unsqueeze(0)
    return x, 1.0 / ratio, 1.0 / ratio


def craft_forward(sd, x, taps=None):
    """VGG_UNet.forward (model.py:39-61) over vgg16_bn.forward (modules/vgg_bn.py:69-82).
    x [N,3,H,W] fp32 -> y [N,H/2,W/2,2].  If `taps` is a dict it receives named intermediates (NCHW)."""
    b = "basenet."

    def cbr(prefix, bn, t, relu=True, **kw):
        t = _bn(sd, bn, _conv(sd, prefix, t, **kw))
        return F.relu(t) if relu else t

    h = cbr(b + "slice1.0", b + "slice1.1", x, padding=1)
    if taps is not None:
        taps["slice1.0"] = h
    h = cbr(b + "slice1.3", b + "slice1.4", h, padding=1)
    h = F.max_pool2d(h, 2, 2)
    h = cbr(b + "slice1.7", b + "slice1.8", h, padding=1)
    # slice1 ends on the BN; slice2 opens with an in-place ReLU on the same storage, so the tap is rectified
    relu2_2 = cbr(b + "slice1.10", b + "slice1.11", h, padding=1)
    h = F.max_pool2d(relu2_2, 2, 2)
    h = cbr(b + "slice2.14", b + "slice2.15", h, padding=1)
    relu3_2 = cbr(b + "slice2.17", b + "slice2.18", h, padding=1)
    h = cbr(b + "slice3.20", b + "slice3.21", relu3_2, padding=1)
    h = F.max_pool2d(h, 2, 2)
    h = cbr(b + "slice3.24", b + "slice3.25", h, padding=1)
    relu4_3 = cbr(b + "slice3.27", b + "slice3.28", h, padding=1)
    h = cbr(b + "slice4.30", b + "slice4.31", relu4_3, padding=1)
    h = F.max_pool2d(h, 2, 2)
    h = cbr(b + "slice4.34", b + "slice4.35", h, padding=1)
    # slice4 also ends on its BN but slice5 starts with a (non in-place) max-pool: this tap keeps its negatives
    relu5_3 = cbr(b + "slice4.37", b + "slice4.38", h, relu=False, padding=1)
    h = F.max_pool2d(relu5_3, 3, 1, 1)
    h = _conv(sd, b + "slice5.1", h, padding=6, dilation=6)
    fc7 = _conv(sd, b + "slice5.2", h)

    def up(prefix, t):
        t = cbr(prefix + ".conv.0", prefix + ".conv.1", t)
        return cbr(prefix + ".conv.3", prefix + ".conv.4", t, padding=1)

    y = up("upconv1", torch.cat([fc7, relu5_3], 1))
    y = F.interpolate(y, size=relu4_3.shape[2:], mode="bilinear", align_corners=False)
    y = up("upconv2", torch.cat([y, relu4_3], 1))
    y = F.interpolate(y, size=relu3_2.shape[2:], mode="bilinear", align_corners=False)
    y = up("upconv3", torch.cat([y, relu3_2], 1))
    y = F.interpolate(y, size=relu2_2.shape[2:], mode="bilinear", align_corners=False)
    feature = up("upconv4", torch.cat([y, relu2_2], 1))
    h = F.relu(_conv(sd, "conv_cls.0", feature, padding=1))
    h = F.relu(_conv(sd, "conv_cls.2", h, padding=1))
    h = F.relu(_conv(sd, "conv_cls.4", h, padding=1))
    h16 = F.relu(_conv(sd, "conv_cls.6", h))
    out = _conv(sd, "conv_cls.8", h16)
    if taps is not None:
        taps.update(relu2_2=relu2_2, relu3_2=relu3_2, relu4_3=relu4_3, relu5_3=relu5_3, fc7=fc7, feature=feature,
                    h16=h16)
    return out.permute(0, 2, 3, 1)


def det_boxes(textmap, linkmap, text_threshold=0.7, link_threshold=0.4, low_text=0.4):
    """det_boxes_core (tools/det_utils.py:35-94): returns (list of float32 [4,2] boxes, labels, kept label ids)."""
    img_h, img_w = textmap.shape
    _, text_score = cv2.threshold(textmap, low_text, 1, 0)
    _, link_score = cv2.threshold(linkmap, link_threshold, 1, 0)
    comb = np.clip(text_score + link_score, 0, 1).astype(np.uint8)
    n, labels, stats, _ = cv2.connectedComponentsWithStats(comb, connectivity=4)
    boxes, kept = [], []
    link_only = np.logical_and(link_score == 1, text_score == 0)
    for k in range(1, n):
        area = stats[k, cv2.CC_STAT_AREA]
        if area < 10:
            continue
        member = labels == k
        if np.max(textmap[member]) < text_threshold:
            continue
        seg = np.zeros(textmap.shape, np.uint8)
        seg[member] = 255
        seg[link_only] = 0
        x, y = stats[k, cv2.CC_STAT_LEFT], stats[k, cv2.CC_STAT_TOP]
        w, h = stats[k, cv2.CC_STAT_WIDTH], stats[k, cv2.CC_STAT_HEIGHT]
        niter = int(math.sqrt(area * min(w, h) / (w * h)) * 2)
        sx, ex, sy, ey = max(x - niter, 0), x + w + niter + 1, max(y - niter, 0), y + h + niter + 1
        ex, ey = min(ex, img_w), min(ey, img_h)
        kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1 + niter, 1 + niter))
        seg[sy:ey, sx:ex] = cv2.dilate(seg[sy:ey, sx:ex], kernel)
        ys, xs = np.nonzero(seg)
        pts = np.stack([xs, ys], 1).reshape(-1, 2)
        box = cv2.boxPoints(cv2.minAreaRect(pts))
        ew, eh = np.linalg.norm(box[0] - box[1]), np.linalg.norm(box[1] - box[2])
        ratio = max(ew, eh) / (min(ew, eh) + 1e-5)
        if abs(1 - ratio) <= 0.1:
            l, r, t, bt = pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max()
            box = np.array([[l, t], [r, t], [r, bt], [l, bt]], np.float32)
        start = box.sum(axis=1).argmin()
        boxes.append(np.array(np.roll(box, 4 - start, 0)))
        kept.append(k)
    return boxes, labels, kept


def rects_from_boxes(boxes, ratio_w, ratio_h, ratio_net=2):
    """adjustResultCoordinates (det_utils.py:259-265) + the rect loop of CRAFT.getCoords (net.py:92-97).
    Returns [min_y, min_x, max_y, max_x] int rects (the reference names them x0, y0, x1, y1)."""
    rects = []
    if len(boxes) == 0:
        return rects
    arr = np.array(boxes)
    for k in range(len(arr)):
        arr[k] *= (ratio_w * ratio_net, ratio_h * ratio_net)
    for box in arr:
        poly = np.array(box).astype(np.int32)
        mn, mx = poly.min(axis=0), poly.max(axis=0)
        rects.append([int(mn[1]), int(mn[0]), int(mx[1]), int(mx[0])])
    return rects


def compare_rects(a, b):
    """Reading-order comparator, branch for branch (det_utils.py:8-26), including its self-comparisons."""
    if a[2] <= b[0]:
        return -1
    if b[2] <= a[0]:
        return 1
    if a[3] <= a[1]:
        return -1
    if b[2] <= b[0]:
        return 1
    for i in (1, 0, 3, 2):
        if a[i] != b[i]:
            return -1 if a[i] < b[i] else 1
    return 0


def sort_rects(rects):
    return sorted(rects, key=cmp_to_key(compare_rects))


def craft_process(sd, image, return_all=False):
    """CRAFT.process (net.py:100-113): image uint8 BGR -> list of crop views in reading order."""
    with torch.no_grad():
        x, rw, rh = craft_preproc(image)
        y = craft_forward(sd, x)
        text = y[0, :, :, 0].numpy().copy()
        link = y[0, :, :, 1].numpy().copy()
    boxes, labels, kept = det_boxes(text, link)
    rects = rects_from_boxes(boxes, rw, rh)
    srt = sort_rects(rects)
    roi = [image[r[0]:r[2], r[1]:r[3], :] for r in srt]
    if return_all:
        return roi, dict(text=text, link=link, boxes=boxes, labels=labels, kept=kept, rects=rects, sorted_rects=srt)
    return roi


# ------------------------------------------------------------------------------------------------ CRNN
def crop_to_tensor(gray):
    """CRNN.getPreds front (net.py:155-158) + ResizeNormalize((100,32)) (tools/dataset.py:43-47).
    gray uint8 [h,w] -> (uint8 [32,100] resized image, fp32 [1,1,32,100])."""
    img = Image.fromarray(gray).convert("L").resize((specs.IMG_W, specs.IMG_H), Image.BICUBIC)
    u8 = np.asarray(img, np.uint8)
    t = torch.from_numpy(u8.astype(np.float32) / np.float32(255.0)).view(1, 1, specs.IMG_H, specs.IMG_W)
    t = t.sub(0.5).div(0.5)
    return u8, t


def bgr_to_gray(bgr):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) (pipeline.py:75)."""
    return cv2.cvtColor(np.ascontiguousarray(bgr), cv2.COLOR_BGR2GRAY)


def tps_localization(sd, x):
    """LocalizationNetwork.forward (modules/TPS_STN.py:70-76): [B,1,32,100] -> fiducials [B,20,2]."""
    p = specs.LOC
    h = x
    for i, (prefix, _, _, _, bn) in enumerate(specs.CRNN_LOC_CONVS):
        h = F.relu(_bn(sd, bn, _conv(sd, prefix, h, padding=1)))
        if i < 3:
            h = F.max_pool2d(h, 2, 2)
    h = F.adaptive_avg_pool2d(h, 1).view(x.shape[0], -1)
    h = F.relu(F.linear(h, sd[p + "localization_fc1.0.weight"], sd[p + "localization_fc1.0.bias"]))
    h = F.linear(h, sd[p + "localization_fc2.weight"], sd[p + "localization_fc2.bias"])
    return h.view(x.shape[0], specs.NUM_FIDUCIAL, 2)


def tps_grid(sd, fid):
    """GridGenerator.build_P_prime (TPS_STN.py:142-150): fiducials [B,20,2] -> sampling grid [B,32,100,2]."""
    B = fid.shape[0]
    inv = sd["Transformation.GridGenerator.inv_delta_C"].unsqueeze(0).expand(B, -1, -1)
    p_hat = sd["Transformation.GridGenerator.P_hat"].unsqueeze(0).expand(B, -1, -1)
    cz = torch.cat([fid, torch.zeros(B, 3, 2)], 1)
    T = torch.bmm(inv, cz)
    return torch.bmm(p_hat, T).reshape(B, specs.IMG_H, specs.IMG_W, 2)


def tps_rectify(sd, x, taps=None):
    """TPS_STN.forward (TPS_STN.py:22-29)."""
    fid = tps_localization(sd, x)
    grid = tps_grid(sd, fid)
    out = F.grid_sample(x, grid, padding_mode="border", align_corners=True)
    if taps is not None:
        taps.update(fiducials=fid, grid=grid, rectified=out)
    return out


def resnet_features(sd, x, taps=None):
    """ResNet.forward (modules/resnet50v1.py:101-135) with BasicBlock.forward (:33-48)."""
    fe = specs.FE

    def cbr(name, bn, t, **kw):
        return F.relu(_bn(sd, fe + bn, _conv(sd, fe + name, t, **kw)))

    def layer(idx, t):
        for i in range(specs.RESNET_BLOCKS[idx]):
            p = "%slayer%d.%d." % (fe, idx, i)
            r = t
            o = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", t, padding=1)))
            o = _bn(sd, p + "bn2", _conv(sd, p + "conv2", o, padding=1))
            if (p + "downsample.0.weight") in sd:
                r = _bn(sd, p + "downsample.1", _conv(sd, p + "downsample.0", t))
            t = F.relu(o + r)
        return t

    h = cbr("conv0_1", "bn0_1", x, padding=1)
    h = cbr("conv0_2", "bn0_2", h, padding=1)
    h = F.max_pool2d(h, 2, 2)
    h = cbr("conv1", "bn1", layer(1, h), padding=1)
    h = F.max_pool2d(h, 2, 2)
    h = cbr("conv2", "bn2", layer(2, h), padding=1)
    h = F.max_pool2d(h, 2, (2, 1), (0, 1))
    h = cbr("conv3", "bn3", layer(3, h), padding=1)
    h = layer(4, h)
    if taps is not None:
        taps["layer4"] = h
    h = cbr("conv4_1", "bn4_1", h, stride=(2, 1), padding=(0, 1))
    h = cbr("conv4_2", "bn4_2", h)
    return h


def _bilstm(sd, prefix, x):
    """BidirectionalLSTM.forward (modules/biLSTM.py:21-33), batch_first, zero initial state, gate order i,f,g,o."""
    n_in = sd[prefix + ".rnn.weight_ih_l0"].shape[1]
    rnn = torch.nn.LSTM(n_in, specs.HIDDEN, bidirectional=True, batch_first=True)
    rnn.load_state_dict({k[len(prefix) + 5:]: v for k, v in sd.items() if k.startswith(prefix + ".rnn.")})
    rnn.eval()
    rec, _ = rnn(x)
    return F.linear(rec, sd[prefix + ".linear.weight"], sd[prefix + ".linear.bias"])


def attention_decode(sd, ctx, num_classes=38, steps=specs.SEQ_T, return_hidden=False):
    """Attention.forward, inference branch (modules/attention.py:46-59) + AttentionCell.forward (:74-88), one crop
    at a time exactly like the reference drives it (batch 1; `h2h(h).unsqueeze(0)` only broadcasts for B=1)."""
    p = "Prediction.attention_cell."
    Hn = specs.HIDDEN
    out = []
    hidden = []
    for b in range(ctx.shape[0]):
        feats = ctx[b:b + 1]                                  # [1,T,256]
        h = torch.zeros(1, Hn)
        c = torch.zeros(1, Hn)
        target = torch.zeros(1, dtype=torch.long)
        probs = torch.zeros(1, steps, num_classes)
        for i in range(steps):
            onehot = torch.zeros(1, num_classes).scatter_(1, target.unsqueeze(1), 1)
            fp = F.linear(feats, sd[p + "i2h.weight"])
            hp = F.linear(h, sd[p + "h2h.weight"], sd[p + "h2h.bias"]).unsqueeze(0)
            e = F.linear(torch.tanh(fp + hp), sd[p + "score.weight"])           # [1,T,1]
            alpha = F.softmax(e, dim=1)
            context = torch.bmm(alpha.permute(0, 2, 1), feats).squeeze(1)       # [1,256]
            xin = torch.cat([context, onehot], 1)
            gates = (F.linear(xin, sd[p + "rnn.weight_ih"], sd[p + "rnn.bias_ih"])
                     + F.linear(h, sd[p + "rnn.weight_hh"], sd[p + "rnn.bias_hh"]))
            gi, gf, gg, go = gates.chunk(4, 1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
            hidden.append(h)
            step = F.linear(h, sd["Prediction.generator.weight"], sd["Prediction.generator.bias"])
            probs[:, i, :] = step
            target = step.max(1)[1]
        out.append(probs)
    if return_hidden:
        return torch.cat(out, 0), torch.cat(hidden, 0)
    return torch.cat(out, 0)


def crnn_forward(sd, x, head="CTC", taps=None):
    """CRNNet.forward (model.py:103-118): x [B,1,32,100] -> preds [B,26,C]."""
    r = tps_rectify(sd, x, taps)
    v = resnet_features(sd, r, taps)                                             # [B,512,1,26]
    v = F.adaptive_avg_pool2d(v.permute(0, 3, 1, 2), (None, 1)).squeeze(3)       # [B,26,512]
    s = _bilstm(sd, "SequenceModeling.0", v)
    s = _bilstm(sd, "SequenceModeling.1", s).contiguous()
    if taps is not None:
        taps.update(visual=v, contextual=s)
    if head == "CTC":
        return F.linear(s, sd["Prediction.weight"], sd["Prediction.bias"])
    if head == "skip":
        return s
    return attention_decode(sd, s, sd["Prediction.generator.weight"].shape[0])


def ctc_decode(idx):
    """CTCLabelConverter.decode for one sequence (tools/recog_utils.py:32-47): drop blanks (0) and repeats."""
    chars = ["[blank]"] + list(specs.ALPHABET)
    out = []
    for i, t in enumerate(idx):
        t = int(t)
        if t != 0 and not (i > 0 and int(idx[i - 1]) == t):
            out.append(chars[t])
    return "".join(out)


def attn_decode_tokens(idx):
    """AttnLabelConverter.decode (recog_utils.py:113-119): all 26 tokens joined, literals '[GO]' / '[s]' included."""
    chars = ["[GO]", "[s]"] + list(specs.ALPHABET)
    return "".join(chars[int(i)] for i in idx)


def crnn_get_preds(sd, gray, head="CTC"):
    """CRNN.getPreds (net.py:152-172): gray uint8 [h,w] -> (raw_pred, preds [1,26,C])."""
    with torch.no_grad():
        _, x = crop_to_tensor(gray)
        preds = crnn_forward(sd, x, head)
    idx = preds.max(2)[1]
    if head == "CTC":
        return [ctc_decode(idx.view(-1))], preds
    return [attn_decode_tokens(idx[0])], preds


def crnn_process(sd, result, gray, head="CTC", verbose=False):
    """CRNN.process (net.py:174-193): mutates and returns `result` {0-d confidence tensor: prediction}."""
    raw_pred, preds = crnn_get_preds(sd, gray, head)
    max_probs = F.softmax(preds, dim=2).max(dim=2)[0]
    for mp in max_probs:
        if head != "CTC":
            pos = raw_pred[0].find("[s]")
            if pos < 0:
                if verbose:
                    print("Not found EOS token, continue.\n(potential error)")
                continue
            raw_pred = raw_pred[0][:pos]
            mp = mp[:pos]
        conf = mp.cumprod(dim=0)[-1]
        if verbose:
            print(f"results: {raw_pred}\tconfidence score: {conf:.4f}\n")
        result[conf] = raw_pred
    return raw_pred, result


def get_text(craft_sd, crnn_sd, image, head="CTC"):
    """getText (pipeline.py:65-87) on an already decoded BGR image; returns the result dict."""
    res = {}
    for crop in craft_process(craft_sd, image):
        _, res = crnn_process(crnn_sd, res, bgr_to_gray(crop), head)
    return res
