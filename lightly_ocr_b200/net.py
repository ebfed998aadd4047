"""Drop-in replacement for the reference's ocr/net.py: the same `CRAFT` / `CRNN` classes, constructor signatures,
attributes, method names, return types and printed lines - backed by liblocr (CUDA, sm_100a) instead of PyTorch/cv2.

Put this package's directory ahead of the reference's `ocr/` on sys.path (or copy this file over ocr/net.py) and
`pipeline.py` / `server.py` run unchanged: `from net import CRAFT, CRNN`.

Differences that are deliberate:
  * `docker=True` meant "run on the CPU" in the reference (pipeline.py:48).  There is no CPU fallback here: the flag is
    accepted and ignored, everything runs on the B200 (net.py:53-57 of the reference).
  * `CRAFT.process` additionally recognises all crops of the image in one batch on the GPU (the pixels never leave
    the device) and parks the results; `CRNN.process(result, gray)` serves a parked result when the gray crop it is
    handed matches (size + CRC), otherwise it runs the crop by itself.  pipeline.py's one-crop-at-a-time loop thereby
    gets batched throughput with identical outputs.  Set LOCR_EAGER=0 to disable.
"""
import os
import zlib
from collections import OrderedDict, deque

import numpy as np
import torch
import yaml

from . import bridge
from .hostops import ALPHABET, AttnLabelConverter, CTCLabelConverter, sort_rects


def _find_ocr_dir():
    env = os.environ.get("LOCR_OCR_DIR")
    if env:
        return env
    import importlib.util
    try:
        spec = importlib.util.find_spec("pipeline")
        if spec and spec.origin and os.path.exists(os.path.join(os.path.dirname(spec.origin), "config.yml")):
            return os.path.dirname(spec.origin)
    except (ImportError, ValueError):
        pass
    return os.path.dirname(os.path.realpath(__file__))


_OCR_DIR = _find_ocr_dir()
DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
MODEL_PATH = os.path.join(_OCR_DIR, "save_models")
with open(os.path.join(_OCR_DIR, "config.yml"), "r") as _yf:
    CONFIG = yaml.safe_load(_yf)


def copyStateDict(state_dict):
    """Strip one leading 'module.' level (reference ocr/net.py:24-34)."""
    start = 1 if list(state_dict.keys())[0].startswith("module") else 0
    out = OrderedDict()
    for k, v in state_dict.items():
        out[".".join(k.split(".")[start:])] = v
    return out


_ENGINES = {}
_PARKED = {}   # (h, w, crc32 of the gray crop) -> deque of parked recognition results
_HELPER = None  # one helper thread for CRAFT._park, started on first use


def _engine(device):
    idx = 0
    if isinstance(device, torch.device) and device.index is not None:
        idx = device.index
    elif isinstance(device, str) and ":" in device:
        idx = int(device.split(":")[1])
    if idx not in _ENGINES:
        act = bridge.ACT_BF16 if os.environ.get("LOCR_ACT", "f16") == "bf16" else bridge.ACT_F16
        head = "CTC" if CONFIG["prediction"] == "CTC" else "Attention"
        _ENGINES[idx] = bridge.Pipeline(device_id=idx, act_dtype=act, head=head, num_classes=CONFIG["num_classes"])
        _ENGINES[idx].crnn_loaded = False
    return _ENGINES[idx]


class Placeholder:
    """Same attribute set as the reference's ABC (ocr/model.py:121-136)."""

    def __init__(self, state_dict=None):
        self.net = None
        self.cuda = False
        self.converter = None
        self.transformer = None
        self.device = None

    def toContainer(self, docker=False):
        pass

    def load(self):
        pass

    def process(self, image):
        pass


class _NetHandle:
    """Stands where the reference keeps an nn.Module (`.net`): callable, with the no-op module methods callers touch."""

    def __init__(self, fn):
        self._fn = fn

    def __call__(self, *a, **kw):
        return self._fn(*a, **kw)

    def eval(self):
        return self

    def to(self, *a, **kw):
        return self

    def parameters(self):
        return iter(())

    def load_state_dict(self, sd, strict=True):
        raise RuntimeError("weights are loaded through CRAFT.load() / CRNN.load()")


_MEAN = np.array([0.485 * 255.0, 0.456 * 255.0, 0.406 * 255.0], np.float32)
_STD = np.array([0.229 * 255.0, 0.224 * 255.0, 0.225 * 255.0], np.float32)


class CRAFT(Placeholder):
    def __init__(self, stateDictName="CRAFT.pth", device=DEVICE, docker=False):
        super().__init__()
        self.model_path = os.path.join(MODEL_PATH, stateDictName)
        self.device = device
        self.engine = _engine(device)
        self.net = _NetHandle(self._forward_tensor)
        if docker:
            self.toContainer(docker=docker)
        self.canvas_size = 1280
        self.magnify_ratio = 1.5
        self.txtThreshold = 0.7
        self.linkThreshold = 0.4
        self.lowTxtScore = 0.4
        self.enablePoly = False
        self.load()

    def toContainer(self, docker=False):
        pass  # no CPU path: the engine always lives on the GPU

    def load(self):
        sd = copyStateDict(torch.load(self.model_path, map_location="cpu"))
        self.engine.load_state_dict(bridge.MODEL_CRAFT, sd)

    def _forward_tensor(self, x):
        """`y, feature = self.net(x)` for tensors produced by preproc(): the uint8 canvas is recovered exactly."""
        arr = x.detach().cpu().numpy()
        canvas = np.rint(arr.transpose(0, 2, 3, 1) * _STD + _MEAN).clip(0, 255).astype(np.uint8)
        y = self.engine.craft_scores(canvas)
        feat = self.engine.debug_read("feature")
        if feat.shape[2] == y.shape[2] + 3:       # row-padded tensor (LOCR_HEAD_HALO=0): one zero pixel left, two right
            feat = feat[:, :, 1:-2]
        feat = feat.transpose(0, 3, 1, 2)
        return torch.from_numpy(y), torch.from_numpy(np.ascontiguousarray(feat))

    def preproc(self, image):
        """Host restatement of the reference's preproc (net.py:71-80) for callers that want the tensor; process()
        itself does this work on the GPU."""
        import cv2
        h, w, ch = image.shape
        target = self.magnify_ratio * max(h, w)
        if target > self.canvas_size:
            target = self.canvas_size
        ratio = target / max(h, w)
        th, tw = int(h * ratio), int(w * ratio)
        proc = cv2.resize(image, (tw, th), interpolation=cv2.INTER_LINEAR)
        h32 = th if th % 32 == 0 else th + (32 - th % 32)
        w32 = tw if tw % 32 == 0 else tw + (32 - tw % 32)
        canvas = np.zeros((h32, w32, ch), np.float32)
        canvas[:th, :tw, :] = proc
        canvas -= _MEAN
        canvas /= _STD
        x = torch.from_numpy(canvas).permute(2, 0, 1).unsqueeze(0)
        return x, 1 / ratio, 1 / ratio

    def getCoords(self, inputs, ratio_w, ratio_h):
        score = np.stack([np.asarray(inputs[0], np.float32), np.asarray(inputs[1], np.float32)], -1)[None]
        out = self.engine.postproc(score, ratio_w, ratio_h, want_labels=False)[0]
        return [[int(v) for v in r] for r in out["rects"]]

    def process(self, image):
        rects_per_image, _, _ = self.engine.detect([image])
        rects = [[int(v) for v in r] for r in rects_per_image[0]]
        srt = sort_rects(rects)
        roi = [image[r[0]:r[2], r[1]:r[3], :] for r in srt]
        if srt and self.engine.crnn_loaded and os.environ.get("LOCR_EAGER", "1") != "0":
            self._park(roi, srt)
        return roi

    def _park(self, roi, srt):
        import cv2
        keep = [i for i, c in enumerate(roi) if c.shape[0] > 0 and c.shape[1] > 0]
        if not keep:
            return
        # the recognition call (ctypes releases the GIL while the GPU works) runs on a helper thread while this one
        # computes the keys the crops will be looked up under: ~0.9 ms of host work per receipt off the critical path
        global _HELPER
        if _HELPER is None:
            from concurrent.futures import ThreadPoolExecutor
            _HELPER = ThreadPoolExecutor(max_workers=1)
        fut = _HELPER.submit(self.engine.recognize_boxes, [0] * len(keep), [srt[i] for i in keep], want_logits=True)
        keys = []
        for i in keep:
            gray = cv2.cvtColor(roi[i], cv2.COLOR_BGR2GRAY)
            keys.append((gray.shape[0], gray.shape[1], zlib.crc32(gray.tobytes())))
        out = fut.result()
        _PARKED.clear()
        for j, key in enumerate(keys):
            _PARKED.setdefault(key, deque()).append({k: (v[j] if v is not None else None) for k, v in out.items()})


class _ResizeNormalize:
    """The reference's transformer attribute (ocr/tools/dataset.py:37-47) for callers that use it directly."""

    def __init__(self, size):
        self.size = size

    def __call__(self, img):
        from PIL import Image
        img = img.resize(self.size, Image.BICUBIC)
        t = torch.from_numpy(np.asarray(img, np.float32) / np.float32(255.0)).unsqueeze(0)
        return t.sub_(0.5).div_(0.5)


class CRNN(Placeholder):
    def __init__(self, stateDictName="CRNN.pth", device=DEVICE, docker=False):
        super().__init__()
        self.alphabet = ALPHABET
        self.model_path = os.path.join(MODEL_PATH, stateDictName)
        self.engine = _engine(device)
        self.net = _NetHandle(self._forward_tensor)
        if docker:
            self.toContainer(docker=docker)
        self.config = CONFIG
        self.device = device
        self.load()

    def toContainer(self, docker=False):
        pass

    def load(self):
        self.transformer = _ResizeNormalize((100, 32))
        sd = torch.load(self.model_path, map_location="cpu")   # the reference does not strip 'module.' here
        self.engine.load_state_dict(bridge.MODEL_CRNN, sd)
        self.engine.crnn_loaded = True
        if self.config["prediction"] == "CTC":
            self.converter = CTCLabelConverter(self.alphabet)
        else:
            self.converter = AttnLabelConverter(self.alphabet)

    def _forward_tensor(self, image, text=None, training=False):
        arr = image.detach().cpu().numpy().reshape(-1, 32, 100)
        u8 = np.rint((arr * 0.5 + 0.5) * 255.0).clip(0, 255).astype(np.uint8)
        return torch.from_numpy(self.engine.crnn_on_resized(u8)["logits"])

    def _run(self, inputs):
        gray = np.ascontiguousarray(inputs)
        key = (gray.shape[0], gray.shape[1], zlib.crc32(gray.tobytes()))
        q = _PARKED.get(key)
        if q:
            return q.popleft()
        out = self.engine.recognize([gray], want_logits=True)
        return {k: (v[0] if v is not None else None) for k, v in out.items()}

    def _raw(self, r):
        if self.config["prediction"] == "CTC":
            return [r["text"]]
        return self.converter.decode(r["ids"][None], [self.config["batch_max_len"]])

    def getPreds(self, inputs):
        r = self._run(inputs)
        return self._raw(r), torch.from_numpy(np.ascontiguousarray(r["logits"]))[None]

    def process(self, result: dict, image: np.ndarray):
        r = self._run(image)
        raw_pred = self._raw(r)
        if self.config["prediction"] == "Attention":
            if r["has_eos"] == 0:
                print("Not found EOS token, continue.\n(potential error)")
                return raw_pred, result
            if r["has_eos"] == -1:
                raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
            raw_pred = r["text"]
        conf = float(r["conf"])
        confidence = torch.tensor(conf, dtype=torch.float32)
        # the reference formats the 0-d tensor (net.py:192), i.e. tensor.item(): the same double as the fp32 value here
        print(f"results: {raw_pred}\tconfidence score: {conf:.4f}\n")
        result[confidence] = raw_pred
        return raw_pred, result
