"""Serving front: the reference's `serveModel` (ocr/pipeline.py:90-112) and the `/api` handler of
ocr/server.py:49-53 with REQUEST BATCHING.  The reference serves one upload at a time through one detector / recogniser
pass per image and per crop; here concurrent requests are collected for a few milliseconds and go through ONE batched
detect + recognise pass of liblocr (`OcrRunner.ocr`), then fan back out to their callers.

    m = serveModel(config_file="config.yml", thresh=0.7, docker=False)      # same constructor as the reference
    m.predict("/path/to/upload.png")      # same result: the predictions whose confidence exceeds `thresh`
    api_response(m, "/path/to/upload.png")  # {'status': 'OK', 'results': {0: ['str'], 1: ...}} like server.py:52-53

`predict` blocks its caller until the batch that contains the request is done, so a threaded WSGI server (or the
`parseText` handler of the reference's server.py, unchanged) gets batching for free.  Results per request are identical
to the reference's one-at-a-time `getText` + threshold filter (tests/test_serve_gpu.py).
"""
import os
import queue
import threading
import time

import cv2
import numpy as np
import torch
import yaml

from . import bridge
from . import net as _net


class _Request:
    __slots__ = ("path", "image", "blob", "done", "result", "error")

    def __init__(self, path):
        self.path = path
        self.image = None
        self.blob = None
        self.done = threading.Event()
        self.result = None
        self.error = None


class serveModel:
    """Drop-in for pipeline.serveModel (same constructor, `predict`, attributes `config`, `thresh`, `model_root`,
    `detector`, `recognizer`), serving through a batching worker."""

    def __init__(self, config_file="config.yml", thresh=0.7, docker=False, max_batch=8, max_wait_ms=4.0, device_id=0):
        self.docker = docker            # accepted like the reference's flag; there is no CPU path
        self.config_file = config_file
        self.loadConfig()
        self.thresh = thresh
        self.model_root = self.config.get("pretrained")
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self.device_id = device_id
        self.batches = []               # sizes of the batches served so far (observability / tests)
        self.encoded_batches = 0        # batches that went through the GPU JPEG decoder without touching host pixels
        self.retried_batches = 0        # batches whose call failed and whose members were then served one at a time
        self.loadModel()
        self._q = queue.Queue()
        self._stop = False
        self._worker = threading.Thread(target=self._loop, daemon=True)
        self._worker.start()

    def loadConfig(self):
        path = self.config_file
        if not os.path.isabs(path):
            path = os.path.join(_net._OCR_DIR, path)
        with open(path, "r") as cf:
            self.config = yaml.safe_load(cf)

    def loadModel(self):
        use_detector, use_recognizer = self.config["pipeline"].split("-")
        if use_detector != "CRAFT":
            raise AssertionError(f"only supported CRAFT atm. got {use_detector} instead")      # pipeline.py:53
        if use_recognizer != "CRNN":
            raise AssertionError(f"only supports either CRNN or MORAN. got {use_recognizer} instead")  # pipeline.py:57
        head = "CTC" if self.config["prediction"] == "CTC" else "Attention"
        self.head = head
        act = bridge.ACT_BF16 if os.environ.get("LOCR_ACT", "f16") == "bf16" else bridge.ACT_F16
        self.runner = bridge.OcrRunner(device_id=self.device_id, act_dtype=act, head=head,
                                       num_classes=self.config["num_classes"])
        craft_sd = _net.copyStateDict(torch.load(os.path.join(_net.MODEL_PATH, "CRAFT.pth"), map_location="cpu"))
        self.runner.load_state_dict(bridge.MODEL_CRAFT, craft_sd)
        self.runner.load_state_dict(bridge.MODEL_CRNN, torch.load(os.path.join(_net.MODEL_PATH, "CRNN.pth"),
                                                                  map_location="cpu"))
        self.detector = self.recognizer = self.runner     # the attributes the reference exposes

    # ------------------------------------------------------------------------------------------------ serving
    def predict(self, inputs: str):
        """pipeline.serveModel.predict: predictions of `inputs` (an image path) with confidence > thresh."""
        req = _Request(inputs)
        self._q.put(req)
        req.done.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def close(self):
        self._stop = True
        self._q.put(None)
        self._worker.join(timeout=5)
        self.runner.close()

    def _loop(self):
        while not self._stop:
            first = self._q.get()
            if first is None:
                break
            batch = [first]
            deadline = time.perf_counter() + self.max_wait
            while len(batch) < self.max_batch:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                try:
                    r = self._q.get(timeout=left)
                except queue.Empty:
                    break
                if r is None:
                    self._stop = True
                    break
                batch.append(r)
            self._serve(batch)

    @staticmethod
    def _read(r):
        """pipeline.py:68 `cv2.imread(path)`.  JPEG uploads (baseline, sequential or progressive Huffman files) and PNG
        uploads stay encoded (r.blob): liblocr decodes them on the GPU, bit-identical to OpenCV incl. the EXIF rotation
        (include/locr.h locr_detect_encoded).  Files outside that subset - arithmetic-coded or CMYK JPEG, animated
        PNG, damaged files ... - and other formats are read by OpenCV itself, exactly like the reference does."""
        with open(r.path, "rb") as f:
            data = f.read()
        if data[:2] == b"\xff\xd8" or data[:8] == b"\x89PNG\r\n\x1a\n":
            try:
                bridge.image_info(data)
                r.blob = data
                return
            except bridge.LocrError:
                pass
        r.image = cv2.imread(r.path)       # not imdecode: only the file reader salvages a truncated JPEG scan
        if r.image is None:
            raise ValueError("cv2.imread could not read %r" % (r.path,))

    def _serve(self, batch):
        good = []
        for r in batch:
            try:
                self._read(r)
                good.append(r)
            except Exception as e:                     # noqa: BLE001 - handed to the caller of predict()
                r.error = ValueError("cv2.imread could not read %r" % (r.path,)) if isinstance(e, OSError) else e
                r.done.set()
        if not good:
            return
        self.batches.append(len(good))
        self._serve_group(good)

    def _run(self, group):
        """One detect + recognise pass over the requests of `group`: (sorted rects per image, recognition outputs)."""
        if all(r.blob is not None for r in group) and hasattr(self.runner, "ocr_encoded"):
            self.encoded_batches += 1
            per_image, out, _ = self.runner.ocr_encoded([r.blob for r in group])
            return per_image, out
        for r in group:                                # mixed batch: the JPEG members are decoded on the GPU as well
            if r.image is None:
                r.image = (self.runner.imdecode(r.blob) if hasattr(self.runner, "imdecode") else
                           cv2.imdecode(np.frombuffer(r.blob, np.uint8), cv2.IMREAD_COLOR))
        return self.runner.ocr([r.image for r in group])

    def _serve_group(self, group):
        """Failures stay with the request that caused them: when a batched call fails (a JPEG whose scan is truncated,
        an image the detector refuses, more boxes than the result buffer holds ...) its members are retried one at a
        time, and a single JPEG that liblocr refuses is handed to OpenCV like any other format - cv2.imread returns a
        partly grey image for a truncated scan and the reference carries on with it (pipeline.py:68)."""
        try:
            per_image, out = self._run(group)
        except Exception as e:                         # noqa: BLE001
            if len(group) > 1:
                self.retried_batches += 1
                for r in group:
                    self._serve_group([r])
                return
            r = group[0]
            if r.blob is not None and isinstance(e, bridge.LocrError):
                img = cv2.imread(r.path)
                r.blob = None
                if img is not None:
                    r.image = img
                    self._serve_group([r])
                    return
                e = ValueError("cv2.imread could not read %r" % (r.path,))
            r.error = e
            r.done.set()
            return
        k = 0
        for r, rects in zip(group, per_image):
            lo, k = k, k + len(rects)
            try:
                r.result = self._filter(out, lo, k)
            except Exception as e:                     # noqa: BLE001 - only this request fails, like in the reference
                r.error = e
            r.done.set()

    def _filter(self, out, lo, hi):
        res = []
        for k in range(lo, hi):
            eos, conf, text = int(out["has_eos"][k]), float(out["conf"][k]), out["text"][k]
            if eos == -2:                              # empty crop: the reference's cv2.cvtColor raises (pipeline.py:75)
                raise ValueError("empty crop")
            if self.head == "CTC":
                value = [text]                         # CRNN.process stores the CTC prediction as a one-element list
            else:
                if eos == 0:                           # no [s]: the reference prints a warning and stores nothing
                    continue
                if eos == -1:
                    raise IndexError("index -1 is out of bounds for dimension 0 with size 0")
                value = text
            # keys of the reference's dict are 0-d float32 tensors: compare in float32 like `k > thresh`
            if np.float32(conf) > self.thresh:
                res.append(value)
        return res


def api_response(model, fpath):
    """Body and status of server.py's POST /api handler (server.py:49-53) for an already saved upload."""
    results = model.predict(fpath)
    return {"status": "OK", "results": {k: v for k, v in enumerate(results)}}, 200
