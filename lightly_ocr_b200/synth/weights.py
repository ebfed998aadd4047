"""Deterministic synthetic checkpoints for CRAFT and CRNN (SURVEY.md 8d weight recipe) - benchmark / test INPUT data.

The reference's weights are Google-Drive downloads (reference scripts/get_model.sh) and cannot be fetched offline, so
both the oracle and the CUDA path load the state dicts produced here; the key names and shapes are the reference's
(oracle/specs.py).  Recipe: biases ~ small, >=2-D weights kaiming-normal (the reference's own training init,
ocr/train/crnn.py:84-97), BatchNorm statistics randomised so that BN folding is exercised, residual branches of the
CRNN ResNet damped (bn2.weight ~ 0.3) and LSTMs at PyTorch's default init to keep rounding-noise amplification of a
random 30-layer network in check (hard part 4).

CRAFT additionally gets an optional "ink path": channel 0 of the layers between the image and the score head carries
a blurred darkness signal, so that synthetic receipts yield word-shaped components (~one box per word) instead of
noise blobs; every other channel stays random.  `calibrate_craft` rescales conv_cls.8 from one forward pass; the calibrated tensors are committed next to this file
(calib_*.npz, produced once by oracle/weights.build_calibrations) so every machine loads identical checkpoints.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import specs


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def _kaiming(g, shape):
    fan_in = int(np.prod(shape[1:]))
    return torch.randn(shape, generator=g) * float(np.sqrt(2.0 / fan_in))


def _bn(sd, g, prefix, c, gamma=None):
    sd[prefix + ".weight"] = torch.rand(c, generator=g) + 0.5 if gamma is None else gamma
    sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
    sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
    sd[prefix + ".running_var"] = torch.rand(c, generator=g) + 0.5
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _bn_identity(sd, prefix, ch=0):
    sd[prefix + ".weight"][ch] = 1.0
    sd[prefix + ".bias"][ch] = 0.0
    sd[prefix + ".running_mean"][ch] = 0.0
    sd[prefix + ".running_var"][ch] = 1.0


def craft_state_dict(seed=0, ink=True):
    """State dict with the reference's VGG_UNet key names (model.py:9-37)."""
    g = _gen(seed)
    sd = OrderedDict()
    for prefix, cin, cout, k, bn, _ in specs.CRAFT_CONVS:
        sd[prefix + ".weight"] = _kaiming(g, (cout, cin, k, k))
        sd[prefix + ".bias"] = torch.randn(cout, generator=g) * 0.05
        if bn is not None:
            _bn(sd, g, bn, cout)
    # keep the un-normalised head small: calibrate_craft sets the final scale
    sd["conv_cls.8.weight"] *= 0.25
    if ink:
        _plant_ink_path(sd)
    return sd


def _plant_ink_path(sd):
    def reserve(prefix, bn, src_ch, taps):
        w = sd[prefix + ".weight"]
        w[0].zero_()
        kh, kw = w.shape[2], w.shape[3]
        if taps == "center":
            w[0, src_ch, kh // 2, kw // 2] = 1.0
        else:  # box blur
            w[0, src_ch] = 1.0 / (kh * kw)
        sd[prefix + ".bias"][0] = 0.0
        if bn is not None:
            _bn_identity(sd, bn)

    # darkness of the normalised BGR pixel: ~1 on ink, 0 on white paper
    w = sd["basenet.slice1.0.weight"]
    w[0].zero_()
    w[0, :, 1, 1] = -0.25 / 3.0
    sd["basenet.slice1.0.bias"][0] = 0.5
    _bn_identity(sd, "basenet.slice1.1")
    reserve("basenet.slice1.3", "basenet.slice1.4", 0, "center")
    reserve("basenet.slice1.7", "basenet.slice1.8", 0, "center")
    reserve("basenet.slice1.10", "basenet.slice1.11", 0, "center")
    # upconv4 input = cat[upsampled y (64 ch), relu2_2 (128 ch)] (model.py:55-57): relu2_2 channel 0 is input 64
    reserve("upconv4.conv.0", "upconv4.conv.1", 64, "center")
    reserve("upconv4.conv.3", "upconv4.conv.4", 0, "box")
    reserve("conv_cls.0", None, 0, "box")
    reserve("conv_cls.2", None, 0, "box")
    reserve("conv_cls.4", None, 0, "box")
    reserve("conv_cls.6", None, 0, "center")


def calibrate_craft(sd, feat16, ink=True):
    """Rescale conv_cls.8 from the 16-channel activations feeding it (feat16: [16, H, W] of one oracle pass).

    Random part: per output channel, median -> 0 and (ink: std -> 0.06 | no ink: 97th percentile -> threshold).
    Ink part: text = 1.6 * z / z_ref, link = 2.4 * z / z_ref where z is channel 0 and z_ref its 99th percentile over
    pixels with any signal, so character cores exceed 0.7 and links bridge the gaps inside a word.
    """
    w = sd["conv_cls.8.weight"].clone()  # [2,16,1,1]
    b = sd["conv_cls.8.bias"].clone()
    f = feat16.reshape(16, -1).double()
    z = f[0]
    for c, thr in ((0, 0.7), (1, 0.4)):
        wr = w[c, :, 0, 0].double().clone()
        if ink:
            wr[0] = 0.0
        r = wr @ f
        med = r.median()
        if ink:
            a = 0.06 / max(float(r.std()), 1e-9)
        else:
            a = thr / max(float(torch.quantile(r[:: max(1, r.numel() // 200000)], 0.97) - med), 1e-9)
        wr = wr * a
        bc = -float(med) * a
        if ink:
            sig = z[z > 0.02]
            z_ref = float(torch.quantile(sig[:: max(1, sig.numel() // 200000)], 0.99)) if sig.numel() else 1.0
            wr[0] = (1.6 if c == 0 else 2.4) / max(z_ref, 1e-9)
        w[c, :, 0, 0] = wr.float()
        b[c] = bc
    sd["conv_cls.8.weight"] = w
    sd["conv_cls.8.bias"] = b
    return sd


_HERE = __import__("os").path.dirname(__import__("os").path.abspath(__file__))


def _overrides(name):
    path = __import__("os").path.join(_HERE, name)
    with np.load(path) as z:
        return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def craft_calibrated(seed=0, ink=True):
    """Raw recipe + the committed calibration (calib_craft_*.npz next to this file):
    BatchNorm running statistics measured on receipt(0) and the rescaled conv_cls.8."""
    assert seed == 0
    sd = craft_state_dict(seed, ink=False)
    ov = _overrides("calib_craft_ink%d.npz" % int(ink))
    for k, v in ov.items():
        if k.startswith("conv_cls.8"):
            continue
        sd[k] = v
    if ink:
        _plant_ink_path(sd)
    sd["conv_cls.8.weight"] = ov["conv_cls.8.weight"]
    sd["conv_cls.8.bias"] = ov["conv_cls.8.bias"]
    return sd


def crnn_calibrated(seed=1, head="CTC", trained=True):
    """Raw recipe + committed calibration (calib_crnn_<head>.npz): BN statistics and the prediction head.

    trained=True (default): the sequence read-out and the early convolutions additionally come from
    calib_crnn_ctc_trained.npz - a short CTC training run with stock PyTorch on synthetic receipts on top of the
    frozen seed-generated front end (tools/train_synth_crnn.py) with the CUDA path's 16-bit storage emulated in the
    forward pass, stored as fp16-representable values - so that the checkpoint decodes confident, input-dependent
    strings like a trained recogniser instead of the near-tie arg-maxes of random weights.
    trained="fp32": the same recipe trained the plain way (fp32 forward pass, no rounding emulation, no injected
    noise, tensors stored unrounded; calib_crnn_*_fp32.npz) - a recogniser that has never seen this repository's
    rounding.
    trained=False: the purely seed-generated checkpoint."""
    assert seed == 1
    sd = crnn_state_dict(seed, head=head)
    sd.update(_overrides("calib_crnn_%s.npz" % head.lower()))
    os_ = __import__("os")
    tag = "fp32" if trained == "fp32" else "trained"
    path = os_.path.join(_HERE, "calib_crnn_ctc_%s.npz" % tag)
    path_a = os_.path.join(_HERE, "calib_crnn_attention_%s.npz" % tag)
    if trained == "fp32" and not (os_.path.exists(path) and (head == "CTC" or os_.path.exists(path_a))):
        raise FileNotFoundError("fp32-trained synthetic checkpoint not generated yet (tools/train_synth_crnn.py)")
    if trained and os_.path.exists(path) and (head == "CTC" or os_.path.exists(path_a)):
        for k, v in _overrides("calib_crnn_ctc_%s.npz" % tag).items():
            if head != "CTC" and k.startswith("Prediction."):
                continue                      # the Attention model shares the trained front end and BiLSTMs only
            assert tuple(v.shape) == tuple(sd[k].shape), k
            sd[k] = v.float()
        if head != "CTC":
            for k, v in _overrides("calib_crnn_attention_%s.npz" % tag).items():
                assert tuple(v.shape) == tuple(sd[k].shape), k
                sd[k] = v.float()
    return sd


def has_fp32_checkpoint(head="CTC"):
    os_ = __import__("os")
    names = ["calib_crnn_ctc_fp32.npz"] + ([] if head == "CTC" else ["calib_crnn_attention_fp32.npz"])
    return all(os_.path.exists(os_.path.join(_HERE, n)) for n in names)


def _pca_head(feats, ncls, gain):
    mean = feats.mean(0)
    u, sv, vt = torch.linalg.svd(feats - mean, full_matrices=False)
    std = sv[:ncls] / np.sqrt(feats.shape[0])
    W = vt[:ncls] * (gain / std).unsqueeze(1)
    return W, -(W @ mean)


# ---------------------------------------------------------------------------------------------- CRNN
def _tps_buffers(F=specs.NUM_FIDUCIAL, H=specs.IMG_H, W=specs.IMG_W, eps=1e-6):
    """inv_delta_C [F+3,F+3] and P_hat [H*W,F+3] of the RARE grid generator (TPS_STN.py:96-140), float64 -> float32."""
    half = F // 2
    cx = np.linspace(-1.0, 1.0, half)
    C = np.concatenate([np.stack([cx, -np.ones(half)], 1), np.stack([cx, np.ones(half)], 1)], 0)  # F x 2
    d = np.linalg.norm(C[:, None, :] - C[None, :, :], axis=2)
    np.fill_diagonal(d, 1.0)
    hat = d ** 2 * np.log(d)
    delta = np.zeros((F + 3, F + 3))
    delta[:F, 0] = 1.0
    delta[:F, 1:3] = C
    delta[:F, 3:] = hat
    delta[F:F + 2, 3:] = C.T
    delta[F + 2, 3:] = 1.0
    inv_delta = np.linalg.inv(delta)
    gx = (np.arange(-W, W, 2) + 1.0) / W
    gy = (np.arange(-H, H, 2) + 1.0) / H
    P = np.stack(np.meshgrid(gx, gy), axis=2).reshape(-1, 2)  # row-major over (y, x)
    rn = np.linalg.norm(P[:, None, :] - C[None, :, :], axis=2)
    rbf = np.square(rn) * np.log(rn + eps)
    P_hat = np.concatenate([np.ones((P.shape[0], 1)), P, rbf], 1)
    return torch.from_numpy(inv_delta).float(), torch.from_numpy(P_hat).float()


def fiducial_template(F=specs.NUM_FIDUCIAL):
    """Initial bias of localization_fc2 (TPS_STN.py:61-68)."""
    half = F // 2
    x = np.linspace(-1.0, 1.0, half)
    top = np.stack([x, np.linspace(0.0, -1.0, half)], 1)
    bot = np.stack([x, np.linspace(1.0, 0.0, half)], 1)
    return torch.from_numpy(np.concatenate([top, bot], 0)).float().view(-1)


def _lstm(sd, g, prefix, n_in, hidden):
    k = 1.0 / np.sqrt(hidden)
    for suffix in ("", "_reverse"):
        sd["%s.weight_ih_l0%s" % (prefix, suffix)] = (torch.rand(4 * hidden, n_in, generator=g) * 2 - 1) * k
        sd["%s.weight_hh_l0%s" % (prefix, suffix)] = (torch.rand(4 * hidden, hidden, generator=g) * 2 - 1) * k
        sd["%s.bias_ih_l0%s" % (prefix, suffix)] = (torch.rand(4 * hidden, generator=g) * 2 - 1) * k
        sd["%s.bias_hh_l0%s" % (prefix, suffix)] = (torch.rand(4 * hidden, generator=g) * 2 - 1) * k


def crnn_state_dict(seed=1, head="CTC", num_classes=None, eos_bias=0.0):
    """State dict with the reference's CRNNet key names (model.py:64-101)."""
    g = _gen(seed)
    H = specs.HIDDEN
    if num_classes is None:
        num_classes = 37 if head == "CTC" else 38
    sd = OrderedDict()
    for prefix, cin, cout, (kh, kw), bn in specs.CRNN_LOC_CONVS:
        sd[prefix + ".weight"] = _kaiming(g, (cout, cin, kh, kw))
        _bn(sd, g, bn, cout)
    sd[specs.LOC + "localization_fc1.0.weight"] = _kaiming(g, (256, 512))
    sd[specs.LOC + "localization_fc1.0.bias"] = torch.randn(256, generator=g) * 0.05
    sd[specs.LOC + "localization_fc2.weight"] = torch.randn(2 * specs.NUM_FIDUCIAL, 256, generator=g) * 0.01
    sd[specs.LOC + "localization_fc2.bias"] = fiducial_template()
    inv_delta, p_hat = _tps_buffers()
    sd["Transformation.GridGenerator.inv_delta_C"] = inv_delta
    sd["Transformation.GridGenerator.P_hat"] = p_hat
    for prefix, cin, cout, (kh, kw), bn in specs.CRNN_FE_CONVS:
        sd[prefix + ".weight"] = _kaiming(g, (cout, cin, kh, kw))
        gamma = None
        if bn.endswith(".bn2") and ".layer" in bn:  # damp the residual branch
            gamma = torch.rand(cout, generator=g) * 0.2 + 0.2
        _bn(sd, g, bn, cout, gamma)
    _lstm(sd, g, "SequenceModeling.0.rnn", 512, H)
    sd["SequenceModeling.0.linear.weight"] = _kaiming(g, (H, 2 * H))
    sd["SequenceModeling.0.linear.bias"] = torch.randn(H, generator=g) * 0.05
    _lstm(sd, g, "SequenceModeling.1.rnn", H, H)
    sd["SequenceModeling.1.linear.weight"] = _kaiming(g, (H, 2 * H))
    sd["SequenceModeling.1.linear.bias"] = torch.randn(H, generator=g) * 0.05
    if head == "CTC":
        sd["Prediction.weight"] = _kaiming(g, (num_classes, H)) * 4.0
        sd["Prediction.bias"] = torch.randn(num_classes, generator=g) * 0.05
    else:
        sd["Prediction.generator.weight"] = _kaiming(g, (num_classes, H)) * 4.0
        gb = torch.randn(num_classes, generator=g) * 0.05
        gb[1] += eos_bias  # [s] token: lets EOS appear at varied positions with random weights
        sd["Prediction.generator.bias"] = gb
        sd["Prediction.attention_cell.i2h.weight"] = _kaiming(g, (H, H))
        sd["Prediction.attention_cell.h2h.weight"] = _kaiming(g, (H, H))
        sd["Prediction.attention_cell.h2h.bias"] = torch.randn(H, generator=g) * 0.05
        sd["Prediction.attention_cell.score.weight"] = _kaiming(g, (1, H))
        k = 1.0 / np.sqrt(H)
        sd["Prediction.attention_cell.rnn.weight_ih"] = (torch.rand(4 * H, H + num_classes, generator=g) * 2 - 1) * k
        sd["Prediction.attention_cell.rnn.weight_hh"] = (torch.rand(4 * H, H, generator=g) * 2 - 1) * k
        sd["Prediction.attention_cell.rnn.bias_ih"] = (torch.rand(4 * H, generator=g) * 2 - 1) * k
        sd["Prediction.attention_cell.rnn.bias_hh"] = (torch.rand(4 * H, generator=g) * 2 - 1) * k
    return sd
