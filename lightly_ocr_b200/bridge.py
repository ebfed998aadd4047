"""ctypes binding of include/locr.h (liblocr.so).  No torch types cross this boundary: numpy / raw pointers only.

The library is built in-tree by lightly_ocr_b200/build.py; if it is missing it is built on first use, and if that is
impossible the import fails loudly (there is no CPU fallback for the product path).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblocr.so")

_lib = None


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "B", "H", "W", "Cin", "Cout", "KH", "KW", "dil_h", "dil_w", "pad_h", "pad_w", "stride_h",
        "x_pitch", "y_pitch", "relu", "out_fp32", "act_dtype", "n_tile")]


class LocrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        _lib = C.CDLL(LIB_PATH)
        _lib.locr_version.restype = C.c_char_p
        _lib.locr_last_error.restype = C.c_char_p
        _lib.locr_last_error.argtypes = [C.c_void_p]
        _lib.locr_test_conv.restype = C.c_int
        _lib.locr_test_conv.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        _lib.locr_test_conv_pool.restype = C.c_int
        _lib.locr_test_conv_pool.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        _lib.locr_test_lstm.restype = C.c_int
        _lib.locr_test_lstm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.POINTER(C.c_float), C.c_int]
        _lib.locr_test_jpeg_coefficients.restype = C.c_int
        _lib.locr_test_jpeg_coefficients.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.locr_jpeg_info.restype = C.c_int
        _lib.locr_jpeg_info.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int)]
        _lib.locr_image_info.restype = C.c_int
        _lib.locr_image_info.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.locr_test_eval_loss.restype = C.c_int
        _lib.locr_test_eval_loss.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.locr_test_png_scanlines.restype = C.c_int
        _lib.locr_test_png_scanlines.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    return _lib


def _check(rc, handle=None):
    if rc != 0:
        msg = lib().locr_last_error(handle)
        raise LocrError("liblocr error %d: %s" % (rc, (msg or b"").decode("utf-8", "replace")))


def _fptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def jpeg_info(data):
    """(height, width, components) of a baseline JPEG file from its header (host only)."""
    data = bytes(data)
    h, w, c = C.c_int(), C.c_int(), C.c_int()
    _check(lib().locr_jpeg_info(data, len(data), C.byref(h), C.byref(w), C.byref(c)))
    return h.value, w.value, c.value


FORMAT_JPEG, FORMAT_PNG = 0, 1


def image_info(data):
    """(height, width, components, format) of a JPEG or PNG file from its header (host only); raises LocrError for
    files outside the subset the GPU ingest covers."""
    data = bytes(data)
    h, w, c, f = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib().locr_image_info(data, len(data), C.byref(h), C.byref(w), C.byref(c), C.byref(f)))
    return h.value, w.value, c.value, f.value


def png_scanlines(data):
    """Host half of the PNG reader (chunk walk, CRC, zlib inflate; no GPU): the filtered scanlines as uint8 [n]."""
    data = bytes(data)
    need = C.c_int64()
    L = lib()
    _check(L.locr_test_png_scanlines(data, len(data), None, 0, C.byref(need)))
    out = np.zeros(need.value, np.uint8)
    _check(L.locr_test_png_scanlines(data, len(data), _fptr(out), out.size, C.byref(need)))
    return out


def jpeg_coefficients(data):
    """Host half of the JPEG reader (marker parsing + Huffman entropy decoding; no GPU): list of int16 arrays
    [block rows][blocks per row][64] (natural order) per component, and the geometry dict."""
    data = bytes(data)
    info = np.zeros(19, np.int32)
    L = lib()
    _check(L.locr_test_jpeg_coefficients(data, len(data), None, 0, _fptr(info)))
    shapes = [(int(info[10 + 4 * k]), int(info[9 + 4 * k])) for k in range(int(info[2]))]
    total = sum(a * b * 64 for a, b in shapes)
    out = np.zeros(total, np.int16)
    _check(L.locr_test_jpeg_coefficients(data, len(data), _fptr(out), total, _fptr(info)))
    planes, base = [], 0
    for a, b in shapes:
        planes.append(out[base:base + a * b * 64].reshape(a, b, 64))
        base += a * b * 64
    geo = dict(height=int(info[0]), width=int(info[1]), hmax=int(info[3]), vmax=int(info[4]), mcux=int(info[5]),
               mcuy=int(info[6]), sampling=[(int(info[7 + 4 * k]), int(info[8 + 4 * k])) for k in range(int(info[2]))])
    return planes, geo


def test_conv(x, w, bias=None, residual=None, *, dil=(1, 1), pad=(0, 0), stride_h=1, relu=False, out_fp32=False,
              act_dtype=1, n_tile=0, x_pitch=None, y_pitch=None):
    """x [B,H,W,x_pitch] fp32 NHWC (first Cin channels used), w [Cout,KH,KW,Cin] -> y [B,OH,OW,y_pitch] fp32."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    B, H, W, xp = x.shape
    Cout, KH, KW, Cin = w.shape
    if x_pitch is None:
        x_pitch = xp
    assert x_pitch == xp
    if y_pitch is None:
        y_pitch = Cout
    OH = (H + 2 * pad[0] - dil[0] * (KH - 1) - 1) // stride_h + 1
    OW = (W + 2 * pad[1] - dil[1] * (KW - 1) - 1) + 1
    d = ConvDesc(B, H, W, Cin, Cout, KH, KW, dil[0], dil[1], pad[0], pad[1], stride_h, x_pitch, y_pitch,
                 int(relu), int(out_fp32), act_dtype, n_tile)
    y = np.zeros((B, OH, OW, y_pitch), np.float32)
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32)
    if residual is not None:
        residual = np.ascontiguousarray(residual, np.float32)
    _check(lib().locr_test_conv(C.byref(d), _fptr(x), _fptr(w), _fptr(bias), _fptr(residual), _fptr(y)))
    return y


def test_conv_pool(x, w, bias=None, *, pad=(1, 1), relu=True, act_dtype=0, n_tile=0, want_full=True):
    """conv + ReLU with the fused MaxPool2d(2, 2): returns (y or None, y_pool [B,OH//2,OW//2,Cout])."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    B, H, W, xp = x.shape
    Cout, KH, KW, Cin = w.shape
    OH = H + 2 * pad[0] - (KH - 1)
    OW = W + 2 * pad[1] - (KW - 1)
    d = ConvDesc(B, H, W, Cin, Cout, KH, KW, 1, 1, pad[0], pad[1], 1, xp, Cout, int(relu), 0, act_dtype, n_tile)
    y = np.zeros((B, OH, OW, Cout), np.float32) if want_full else None
    yp = np.zeros((B, OH // 2, OW // 2, Cout), np.float32)
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32)
    _check(lib().locr_test_conv_pool(C.byref(d), _fptr(x), _fptr(w), _fptr(bias), _fptr(y), _fptr(yp)))
    return y, yp


def test_lstm(xproj, whh, act_dtype=0, iters=0, split=0):
    """The BiLSTM recurrence kernel alone: xproj [B,T,2048] fp32 (W_ih x + biases, PyTorch row order), whh [2,1024,256]
    -> hidden states [B,T,512] fp32 (and ms per launch when iters > 0).  split=1: the split-precision feedback of
    LOCR_PREC_EXACT (h carried as hi + lo; the returned states are hi + lo)."""
    xproj = np.ascontiguousarray(xproj, np.float32)
    whh = np.ascontiguousarray(whh, np.float32)
    B, T, n = xproj.shape
    assert n == 2048 and whh.shape == (2, 1024, 256)
    out = np.zeros((B, T, 512), np.float32)
    ms = C.c_float(0)
    _check(lib().locr_test_lstm(_fptr(xproj), _fptr(whh), B, T, act_dtype, _fptr(out), iters, C.byref(ms), int(split)))
    return (out, ms.value) if iters > 0 else out


def test_eval_loss(logits, targets, target_len, head="CTC"):
    """The evaluation-loss kernels alone on host logits [n,26,C].  CTC: targets = concatenated class indices,
    target_len [n] -> (loss [n], correct [n]).  Attention: targets [n, batch_max_len + 2] (AttnLabelConverter.encode
    rows) -> (loss sums [n], counted steps [n], correct [n])."""
    logits = np.ascontiguousarray(logits, np.float32)
    n, T, nc = logits.shape
    assert T == 26
    tg = np.ascontiguousarray(targets, np.int32)
    tl = np.ascontiguousarray(target_len, np.int32)
    loss = np.zeros(n, np.float32)
    count = np.zeros(n, np.int32)
    correct = np.zeros(n, np.int32)
    attn = head != "CTC"
    _check(lib().locr_test_eval_loss(int(attn), _fptr(logits), n, nc, _fptr(tg), _fptr(tl), tg.size, _fptr(loss),
                                     _fptr(count), _fptr(correct)))
    return (loss, count, correct) if attn else (loss, correct)


class Config(C.Structure):
    _fields_ = [("device_id", C.c_int), ("act_dtype", C.c_int), ("head", C.c_int), ("num_classes", C.c_int),
                ("canvas_size", C.c_int), ("mag_ratio", C.c_float), ("text_threshold", C.c_float),
                ("link_threshold", C.c_float), ("low_text", C.c_float), ("crnn_precision", C.c_int)]


MODEL_CRAFT, MODEL_CRNN = 0, 1
HEAD_CTC, HEAD_ATTN = 0, 1
ACT_F16, ACT_BF16 = 0, 1
PREC_FAST, PREC_EXACT = 0, 1      # include/locr.h LOCR_PREC_*: arithmetic of the recogniser
TEXT_STRIDE = 128


def _bind_engine(L):
    if getattr(L, "_engine_bound", False):
        return
    vp = C.c_void_p
    L.locr_create.restype = C.c_int
    L.locr_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.locr_destroy.restype = None
    L.locr_destroy.argtypes = [vp]
    L.locr_load_tensor.restype = C.c_int
    L.locr_load_tensor.argtypes = [vp, C.c_int, C.c_char_p, vp, C.POINTER(C.c_int64), C.c_int]
    L.locr_finalize.restype = C.c_int
    L.locr_finalize.argtypes = [vp, C.c_int]
    L.locr_launch_count.restype = C.c_int64
    L.locr_launch_count.argtypes = [vp]
    L.locr_debug_craft_scores.restype = C.c_int
    L.locr_debug_craft_scores.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.locr_debug_crnn.restype = C.c_int
    L.locr_debug_crnn.argtypes = [vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp]
    L.locr_audit.restype = C.c_int
    L.locr_audit.argtypes = [vp, C.c_int]
    L.locr_audit_read.restype = C.c_int
    L.locr_audit_read.argtypes = [vp, C.c_char_p, C.c_int64]
    L.locr_debug_read.restype = C.c_int
    L.locr_debug_read.argtypes = [vp, C.c_char_p, vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L._engine_bound = True


class Engine:
    """One liblocr handle = one GPU.  Weights go in as a {name: array-like fp32} mapping with reference key names."""

    def __init__(self, device_id=0, act_dtype=ACT_BF16, head="CTC", num_classes=None, canvas_size=1280,
                 mag_ratio=1.5, text_threshold=0.7, link_threshold=0.4, low_text=0.4, precision=None):
        self.L = lib()
        _bind_engine(self.L)
        self.head = HEAD_CTC if head == "CTC" else HEAD_ATTN
        self.num_classes = num_classes or (37 if self.head == HEAD_CTC else 38)
        if precision is None:
            precision = PREC_EXACT if os.environ.get("LOCR_CRNN_PREC", "fast") == "exact" else PREC_FAST
        self.precision = precision
        cfg = Config(device_id, act_dtype, self.head, self.num_classes, canvas_size, mag_ratio, text_threshold,
                     link_threshold, low_text, precision)
        self.h = C.c_void_p()
        _check(self.L.locr_create(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.locr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, model, state_dict):
        for key, val in state_dict.items():
            a = val.detach().cpu().numpy() if hasattr(val, "detach") else np.asarray(val)
            if a.dtype.kind not in "fiu" or key.endswith("num_batches_tracked"):
                continue
            a = np.ascontiguousarray(a, np.float32)
            shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            _check(self.L.locr_load_tensor(self.h, model, key.encode(), _fptr(a), shape, a.ndim), self.h)
        _check(self.L.locr_finalize(self.h, model), self.h)

    def launch_count(self):
        return int(self.L.locr_launch_count(self.h))

    def audit(self, enable=True):
        """Range audit of the 16-bit activations (include/locr.h locr_audit): enable, run forward passes, audit_read()."""
        _check(self.L.locr_audit(self.h, int(enable)), self.h)

    def audit_read(self):
        """[(layer name, largest |activation| stored)] per convolution launch since audit(True)."""
        buf = C.create_string_buffer(1 << 16)
        _check(self.L.locr_audit_read(self.h, buf, len(buf)), self.h)
        rows = []
        for line in buf.value.decode().splitlines():
            name, v = line.rsplit(" ", 1)
            rows.append((name, float(v)))
        return rows

    def craft_scores(self, images):
        """images: uint8 [B,H,W,3] BGR (same size) -> fp32 [B,H32/2,W32/2,2]."""
        images = np.ascontiguousarray(images, np.uint8)
        B, H, W, _ = images.shape
        H32, W32 = (H + 31) // 32 * 32, (W + 31) // 32 * 32
        out = np.empty((B, H32 // 2, W32 // 2, 2), np.float32)
        _check(self.L.locr_debug_craft_scores(self.h, _fptr(images), B, H, W, _fptr(out)), self.h)
        return out

    def crnn_on_resized(self, u8):
        """u8: uint8 [n,32,100] -> dict(logits, ids, text, has_eos, conf)."""
        u8 = np.ascontiguousarray(u8, np.uint8)
        n = u8.shape[0]
        logits = np.empty((n, 26, self.num_classes), np.float32)
        ids = np.empty((n, 26), np.int32)
        text = np.zeros((n, TEXT_STRIDE), np.uint8)
        eos = np.empty(n, np.int32)
        conf = np.empty(n, np.float32)
        _check(self.L.locr_debug_crnn(self.h, _fptr(u8), n, _fptr(logits), _fptr(ids), _fptr(text), TEXT_STRIDE,
                                      _fptr(eos), _fptr(conf)), self.h)
        strings = [bytes(row).split(b"\0", 1)[0].decode("ascii") for row in text]
        return dict(logits=logits, ids=ids, text=strings, has_eos=eos, conf=conf)

    def debug_read(self, name):
        shape = (C.c_int64 * 8)()
        nd = C.c_int()
        _check(self.L.locr_debug_read(self.h, name.encode(), None, 0, shape, C.byref(nd)), self.h)
        shp = tuple(int(shape[i]) for i in range(nd.value))
        out = np.empty(shp, np.float32)
        _check(self.L.locr_debug_read(self.h, name.encode(), _fptr(out), out.size, shape, C.byref(nd)), self.h)
        return out


def _bind_pipeline(L):
    if getattr(L, "_pipeline_bound", False):
        return
    vp = C.c_void_p
    L.locr_detect.restype = C.c_int
    L.locr_detect.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]
    L.locr_recognize.restype = C.c_int
    L.locr_recognize.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.locr_evaluate.restype = C.c_int
    L.locr_evaluate.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int64, vp, vp, vp, vp, vp,
                                C.POINTER(C.c_float)]
    L.locr_recognize_boxes.restype = C.c_int
    L.locr_recognize_boxes.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.locr_debug_postproc.restype = C.c_int
    L.locr_debug_postproc.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, vp, vp, vp,
                                      vp, vp]
    L.locr_get_det_boxes.restype = C.c_int
    L.locr_get_det_boxes.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                     vp, vp, vp, vp]
    L.locr_debug_resize.restype = C.c_int
    L.locr_debug_resize.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
    L.locr_imdecode.restype = C.c_int
    L.locr_imdecode.argtypes = [vp, C.c_char_p, C.c_int64, vp, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.locr_detect_encoded.restype = C.c_int
    L.locr_detect_encoded.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    L._pipeline_bound = True


def _ptr_array(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def _int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def _texts(buf):
    return [bytes(row).split(b"\0", 1)[0].decode("ascii") for row in buf]


class Pipeline(Engine):
    """Engine + the product entry points (locr_detect / locr_recognize / locr_recognize_boxes)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        _bind_pipeline(self.L)

    def detect(self, images, max_boxes_total=None, want_boxes=False, want_scores=False):
        """images: list of uint8 BGR HxWx3 arrays.  Returns (rects list per image, boxes per image or None,
        score maps per image or None); rects are [min_y, min_x, max_y, max_x] in label order (unsorted)."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        cap = max_boxes_total or 4096 * n
        rects = np.empty((cap, 4), np.int32)
        boxes = np.empty((cap, 4, 2), np.float32) if want_boxes else None
        counts = np.zeros(n, np.int32)
        scores = None
        sizes = []
        if want_scores:
            total = 0
            for im in imgs:
                h, w = im.shape[:2]
                mx = max(h, w)
                target = min(1.5 * mx, 1280)
                ratio = target / mx
                th, tw = int(h * ratio), int(w * ratio)
                h32 = th if th % 32 == 0 else th + (32 - th % 32)
                w32 = tw if tw % 32 == 0 else tw + (32 - tw % 32)
                sizes.append((h32 // 2, w32 // 2))
                total += (h32 // 2) * (w32 // 2) * 2
            scores = np.empty(total, np.float32)
        _check(self.L.locr_detect(self.h, _ptr_array(imgs), _int_array([i.shape[0] for i in imgs]),
                                  _int_array([i.shape[1] for i in imgs]), None, n, cap, _fptr(rects), _fptr(boxes),
                                  _fptr(counts), _fptr(scores)), self.h)
        out_r, out_b, out_s = [], [], []
        base = 0
        sbase = 0
        for i in range(n):
            k = int(counts[i])
            out_r.append(rects[base:base + k].copy())
            if want_boxes:
                out_b.append(boxes[base:base + k].copy())
            if want_scores:
                mh, mw = sizes[i]
                out_s.append(scores[sbase:sbase + mh * mw * 2].reshape(mh, mw, 2).copy())
                sbase += mh * mw * 2
            base += k
        return out_r, (out_b if want_boxes else None), (out_s if want_scores else None)

    def _outputs(self, n, want_logits):
        return dict(logits=np.empty((n, 26, self.num_classes), np.float32) if want_logits else None,
                    ids=np.empty((n, 26), np.int32), text=np.zeros((n, TEXT_STRIDE), np.uint8),
                    has_eos=np.empty(n, np.int32), conf=np.empty(n, np.float32))

    def recognize(self, crops, want_logits=True):
        """crops: list of uint8 arrays, HxW (gray) or HxWx3 (BGR)."""
        cs = [np.ascontiguousarray(c, np.uint8) for c in crops]
        n = len(cs)
        o = self._outputs(n, want_logits)
        ch = [1 if c.ndim == 2 else c.shape[2] for c in cs]
        _check(self.L.locr_recognize(self.h, _ptr_array(cs), _int_array([c.shape[0] for c in cs]),
                                     _int_array([c.shape[1] for c in cs]), None, _int_array(ch), n,
                                     _fptr(o["logits"]), _fptr(o["ids"]), _fptr(o["text"]), _fptr(o["has_eos"]),
                                     _fptr(o["conf"])), self.h)
        o["text"] = _texts(o["text"])
        return o

    def evaluate(self, crops, targets, target_len):
        """One validation batch of the reference's evaluation() (ocr/train/crnn.py:142-240): recognition + the loss and
        the label == prediction flags on the GPU.  targets / target_len as the reference's converter.encode returns them
        (CTC: concatenated indices + lengths; Attention: [n, batch_max_len + 2] rows + lengths).  Returns a dict with
        cost (the scalar loss_fn returns), loss [n], correct [n], ids, text, conf."""
        cs = [np.ascontiguousarray(c, np.uint8) for c in crops]
        n = len(cs)
        tg = np.ascontiguousarray(targets, np.int32)
        tl = np.ascontiguousarray(target_len, np.int32)
        assert tl.size == n
        o = dict(loss=np.zeros(n, np.float32), correct=np.zeros(n, np.int32), ids=np.empty((n, 26), np.int32),
                 text=np.zeros((n, TEXT_STRIDE), np.uint8), conf=np.empty(n, np.float32))
        ch = [1 if c.ndim == 2 else c.shape[2] for c in cs]
        cost = C.c_float(0)
        _check(self.L.locr_evaluate(self.h, _ptr_array(cs), _int_array([c.shape[0] for c in cs]),
                                    _int_array([c.shape[1] for c in cs]), None, _int_array(ch), n, _fptr(tg), _fptr(tl),
                                    tg.size, _fptr(o["loss"]), _fptr(o["correct"]), _fptr(o["ids"]), _fptr(o["text"]),
                                    _fptr(o["conf"]), C.byref(cost)), self.h)
        o["text"] = _texts(o["text"])
        o["cost"] = cost.value
        return o

    def recognize_boxes(self, image_index, rects, want_logits=False, want_u8=False):
        idx = np.ascontiguousarray(image_index, np.int32)
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        n = len(idx)
        o = self._outputs(n, want_logits)
        u8 = np.empty((n, 32, 100), np.uint8) if want_u8 else None
        _check(self.L.locr_recognize_boxes(self.h, _fptr(idx), _fptr(r), n, _fptr(o["logits"]), _fptr(o["ids"]),
                                           _fptr(o["text"]), _fptr(o["has_eos"]), _fptr(o["conf"]), _fptr(u8)),
               self.h)
        o["text"] = _texts(o["text"])
        o["u8"] = u8
        return o

    def imdecode(self, data):
        """cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR) for a JPEG or PNG file, decoded on the GPU
        (entropy decoding / inflate on the host, the rest in CUDA): uint8 [H][W][3] BGR.  Raises LocrError for files
        outside the covered subset (arithmetic-coded or CMYK JPEG, animated PNG, truncated files ...)."""
        data = bytes(data)
        h, w, _, _ = image_info(data)
        out = np.empty((h, w, 3), np.uint8)
        hh, ww = C.c_int(), C.c_int()
        _check(self.L.locr_imdecode(self.h, data, len(data), _fptr(out), out.nbytes, C.byref(hh), C.byref(ww)), self.h)
        return out

    def imread(self, path):
        """cv2.imread(path) for JPEG and PNG files (reference ocr/pipeline.py:68)."""
        with open(path, "rb") as f:
            return self.imdecode(f.read())

    def postproc(self, score, ratio_w=1.0, ratio_h=1.0, max_boxes=4096, want_labels=True):
        """score: fp32 [B,H,W,2] -> per image (boxes [k,4,2], rects [k,4], box labels [k], n components, labels)."""
        score = np.ascontiguousarray(score, np.float32)
        B, H, W, _ = score.shape
        boxes = np.empty((B, max_boxes, 4, 2), np.float32)
        rects = np.empty((B, max_boxes, 4), np.int32)
        lab = np.empty((B, max_boxes), np.int32)
        counts = np.empty((B, 2), np.int32)
        labels = np.empty((B, H, W), np.int32) if want_labels else None
        _check(self.L.locr_debug_postproc(self.h, _fptr(score), B, H, W, ratio_w, ratio_h, max_boxes, _fptr(boxes),
                                          _fptr(rects), _fptr(lab), _fptr(counts), _fptr(labels)), self.h)
        out = []
        for b in range(B):
            k = int(counts[b, 0])
            out.append(dict(boxes=boxes[b, :k].copy(), rects=rects[b, :k].copy(), box_label=lab[b, :k].copy(),
                            n_components=int(counts[b, 1]), labels=None if labels is None else labels[b]))
        return out

    def get_det_boxes(self, textmap, linkmap, text_threshold=0.7, link_threshold=0.4, low_text=0.4, poly=False,
                      max_boxes=4096):
        """tools.getDetBoxes of the reference (ocr/tools/det_utils.py:248-256): (boxes, polys) for one pair of score
        maps - boxes as a list of float32 [4, 2] arrays in score-map coordinates, polys as a list holding a float64
        [14, 2] array or None per box (all None when poly is False)."""
        score = np.ascontiguousarray(np.stack([np.asarray(textmap, np.float32), np.asarray(linkmap, np.float32)], -1))
        H, W, _ = score.shape
        boxes = np.empty((max_boxes, 4, 2), np.float32)
        counts = np.zeros(1, np.int32)
        polys = np.empty((max_boxes, 14, 2), np.float64) if poly else None
        valid = np.zeros(max_boxes, np.int32) if poly else None
        _check(self.L.locr_get_det_boxes(self.h, _fptr(score), 1, H, W, text_threshold, link_threshold, low_text,
                                         int(bool(poly)), max_boxes, _fptr(boxes), _fptr(counts), _fptr(polys),
                                         _fptr(valid)), self.h)
        n = int(counts[0])
        out_b = [boxes[i].copy() for i in range(n)]
        out_p = [polys[i].copy() if (poly and valid[i]) else None for i in range(n)]
        return out_b, out_p

    def resize_linear(self, img, out_w, out_h):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.empty((out_h, out_w, 3), np.uint8)
        _check(self.L.locr_debug_resize(self.h, _fptr(img), img.shape[0], img.shape[1], _fptr(out), out_h, out_w),
               self.h)
        return out


def _bind_bench(L):
    if getattr(L, "_bench_bound", False):
        return
    vp = C.c_void_p
    L.locr_detect_resident.restype = C.c_int
    L.locr_detect_resident.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.locr_timer_start.restype = C.c_int
    L.locr_timer_start.argtypes = [vp]
    L.locr_timer_stop.restype = C.c_int
    L.locr_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.locr_profile.restype = C.c_int
    L.locr_profile.argtypes = [vp, C.c_int]
    L.locr_profile_read.restype = C.c_int
    L.locr_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.locr_profile_layers.restype = C.c_int
    L.locr_profile_layers.argtypes = [vp, C.c_char_p, C.c_int64]
    L._bench_bound = True


class OcrRunner(Pipeline):
    """getText for batches of receipts (reference ocr/pipeline.py:65-87): detect on the GPU, reading-order sort on the
    host exactly like the reference, crops + recognition on the GPU.  Used by bench.py and the throughput tests."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        _bind_bench(self.L)
        self._cap = 0
        self._sorter = None      # helper thread of _sort_and_recognize, started on first use

    def close(self):
        if getattr(self, "_sorter", None) is not None:
            self._sorter.shutdown(wait=True)
            self._sorter = None
        super().close()

    def _buffers(self, n):
        cap = 4096 * n          # the detector keeps up to 4096 boxes per image (pipeline.cu)
        if cap > self._cap:
            self._rects = np.empty((cap, 4), np.int32)
            self._counts = np.zeros(max(n, 1), np.int32)
            self._cap = cap
        if len(self._counts) < n:
            self._counts = np.zeros(n, np.int32)
        return self._rects, self._counts

    def _sort_and_recognize(self, n, rects, counts, want_logits=False):
        """Reading-order sort (net.py:108) + recognition of every crop.  What a crop reads does not depend on its
        position in the batch, so the kernels run on the rects in label order while a helper thread does the sort - a
        Python comparator by contract (hostops.compare_rects), ~1.2 ms per 8 receipts during which the GPU used to idle;
        the ctypes call releases the GIL - and the outputs are permuted afterwards.  LOCR_SORT_OVERLAP=0: sort first."""
        from .hostops import sort_rects
        total = int(np.sum(counts[:n]))
        per_len = [int(counts[i]) for i in range(n)]
        if total == 0:
            return [[] for _ in range(n)], dict(text=[], conf=np.zeros(0, np.float32), has_eos=np.zeros(0, np.int32))
        if os.environ.get("LOCR_SORT_OVERLAP", "1") == "0":
            idx, srt, per_image, base = [], [], [], 0
            for i in range(n):
                s = sort_rects(rects[base:base + per_len[i]].tolist())
                per_image.append(s)
                srt.extend(s)
                idx.extend([i] * per_len[i])
                base += per_len[i]
            return per_image, self.recognize_boxes(idx, srt, want_logits=want_logits)
        flat = rects[:total].tolist()
        idx = np.repeat(np.arange(n, dtype=np.int32), per_len)

        def orders():
            # the comparator only reads elements 0..3: a fifth element carries the position through the sort
            perm, base = [], 0
            for k in per_len:
                tagged = [flat[base + j] + [base + j] for j in range(k)]
                perm.extend(t[4] for t in sort_rects(tagged))
                base += k
            return perm

        if self._sorter is None:
            from concurrent.futures import ThreadPoolExecutor
            self._sorter = ThreadPoolExecutor(max_workers=1)
        fut = self._sorter.submit(orders)
        out = self.recognize_boxes(idx, rects[:total], want_logits=want_logits)
        perm = fut.result()
        per_image, base = [], 0
        for k in per_len:
            per_image.append([flat[q] for q in perm[base:base + k]])
            base += k
        pa = np.asarray(perm, np.int64)
        for key, v in list(out.items()):
            if v is None:
                continue
            out[key] = [v[q] for q in perm] if isinstance(v, list) else v[pa]
        return per_image, out

    def ocr(self, images, want_logits=False):
        """Host images in, (sorted rects per image, recognition outputs over all crops) out."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        rects, counts = self._buffers(n)
        _check(self.L.locr_detect(self.h, _ptr_array(imgs), _int_array([i.shape[0] for i in imgs]),
                                  _int_array([i.shape[1] for i in imgs]), None, n, self._cap, _fptr(rects), None,
                                  _fptr(counts), None), self.h)
        return self._sort_and_recognize(n, rects, counts, want_logits)

    def ocr_encoded(self, blobs, want_logits=False):
        """getText on encoded (baseline JPEG) files: decoded on the GPU straight into the resident image buffer, so the
        pixels never visit the host.  Returns (sorted rects per image, recognition outputs, [(height, width)])."""
        blobs = [bytes(b) for b in blobs]
        n = len(blobs)
        rects, counts = self._buffers(n)
        ptrs = (C.c_char_p * n)(*blobs)
        sizes = (C.c_int64 * n)(*[len(b) for b in blobs])
        hs, ws = (C.c_int * n)(), (C.c_int * n)()
        _check(self.L.locr_detect_encoded(self.h, ptrs, sizes, n, self._cap, _fptr(rects), None, _fptr(counts), None,
                                          hs, ws), self.h)
        per_image, out = self._sort_and_recognize(n, rects, counts, want_logits)
        return per_image, out, [(int(hs[i]), int(ws[i])) for i in range(n)]

    def ocr_resident(self, n, want_logits=False):
        """Same, on the images the previous ocr()/detect() call left resident in HBM (no host-to-device copy)."""
        rects, counts = self._buffers(n)
        _check(self.L.locr_detect_resident(self.h, self._cap, _fptr(rects), None, _fptr(counts), None), self.h)
        return self._sort_and_recognize(n, rects, counts, want_logits)

    def detect_resident(self, n):
        """Detection only (CRAFT forward + boxes) on the n images left resident in HBM: rects per image."""
        rects, counts = self._buffers(n)
        _check(self.L.locr_detect_resident(self.h, self._cap, _fptr(rects), None, _fptr(counts), None), self.h)
        out, base = [], 0
        for i in range(n):
            out.append(rects[base:base + int(counts[i])].copy())
            base += int(counts[i])
        return out

    def timer_start(self):
        _check(self.L.locr_timer_start(self.h), self.h)

    def timer_stop(self):
        ms = C.c_float()
        _check(self.L.locr_timer_stop(self.h, C.byref(ms)), self.h)
        return float(ms.value)

    def profile(self, enable):
        _check(self.L.locr_profile(self.h, int(enable)), self.h)

    def profile_read(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        _check(self.L.locr_profile_read(self.h, C.byref(ms), C.byref(fl), C.byref(n)), self.h)
        return float(ms.value), float(fl.value), int(n.value)

    def profile_layers(self):
        """[(name, ms, flops, launches)] accumulated by the profile_read() calls since the last call."""
        buf = C.create_string_buffer(1 << 16)
        _check(self.L.locr_profile_layers(self.h, buf, len(buf)), self.h)
        rows = []
        for line in buf.value.decode().splitlines():
            name, ms, fl, n = line.rsplit(" ", 3)
            rows.append((name, float(ms), float(fl), int(n)))
        return rows
