"""Serving front: `lightly_ocr_b200.serve.serveModel` batches concurrent `predict` calls; every
request must get exactly what the reference's one-image-at-a-time path returns: getText (ocr/pipeline.py:65-87, driven
here through the drop-in net.CRAFT / net.CRNN classes) followed by serveModel.predict's `k > thresh` filter
(pipeline.py:106-112) and server.py's JSON shape (:52-53)."""
import contextlib
import importlib
import io
import os
import threading

import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(tmp_path):
    from lightly_ocr_b200.synth import weights
    d = tmp_path / "ocr"
    (d / "save_models").mkdir(parents=True)
    torch.save(weights.craft_calibrated(0, ink=True), str(d / "save_models" / "CRAFT.pth"))
    torch.save(weights.crnn_calibrated(1, "CTC"), str(d / "save_models" / "CRNN.pth"))
    import yaml
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "..", "lightly_ocr_b200", "config.yml")))
    cfg["prediction"] = "CTC"
    cfg["num_classes"] = 37
    yaml.safe_dump(cfg, open(str(d / "config.yml"), "w"))
    os.environ["LOCR_OCR_DIR"] = str(d)
    import lightly_ocr_b200.net as net
    for e in getattr(net, "_ENGINES", {}).values():
        e.close()
    net = importlib.reload(net)
    import lightly_ocr_b200.serve as serve
    return net, importlib.reload(serve), d


def test_batched_serving_equals_one_at_a_time(tmp_path):
    from lightly_ocr_b200.synth import receipts
    net, serve, d = _setup(tmp_path)
    paths = []
    for i in range(12):
        img = receipts.receipt(40 + i)
        if i % 3 == 1:
            img = np.ascontiguousarray(img[:640, :480])     # mixed sizes in one batch
        p = str(d / ("upload%d.png" % i))
        cv2.imwrite(p, img)
        paths.append(p)
    thresh = 0.7
    # --- the reference's serial path: getText's loop + predict's filter, on the drop-in classes
    detector, recognizer = net.CRAFT(device=net.DEVICE), net.CRNN(device=net.DEVICE)
    want = []
    for p in paths:
        res = {}
        image = cv2.imread(p)
        with contextlib.redirect_stdout(io.StringIO()):
            for crop in detector.process(image):
                _, res = recognizer.process(res, cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY))
        want.append([v for k, v in res.items() if k > thresh])
    # --- 12 concurrent requests against the batching front
    m = serve.serveModel(config_file="config.yml", thresh=thresh, docker=False, max_batch=8, max_wait_ms=50)
    got = [None] * len(paths)

    def call(i):
        got[i] = serve.api_response(m, paths[i])

    threads = [threading.Thread(target=call, args=(i,)) for i in range(len(paths))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    for i in range(len(paths)):
        body, status = got[i]
        assert status == 200 and body["status"] == "OK"
        assert [body["results"][k] for k in range(len(body["results"]))] == want[i], i
    assert max(m.batches) > 1 and sum(m.batches) == len(paths), m.batches
    kept = sum(len(w) for w in want)
    print("12 concurrent requests served in batches %s; %d predictions above the %.1f threshold" %
          (m.batches, kept, thresh))
    assert kept > 200
    # --- errors stay with their request
    with pytest.raises(ValueError):
        m.predict(str(d / "missing.png"))
    assert m.predict(paths[0]) == want[0]
    # --- JPEG uploads: decoded on the GPU (locr_detect_encoded), same answer as cv2.imread + the serial path
    jpgs = []
    for i in range(3):
        p = str(d / ("photo%d.jpg" % i))
        cv2.imwrite(p, receipts.receipt(60 + i), [cv2.IMWRITE_JPEG_QUALITY, 90])
        jpgs.append(p)
    pp = str(d / "actually_png.jpg")
    with open(pp, "wb") as f:                 # PNG content under a .jpg name: told apart by the signature, GPU route as well
        f.write(cv2.imencode(".png", receipts.receipt(63)[:640, :480])[1].tobytes())
    jpgs.append(pp)
    bm = str(d / "scan.bmp")                  # a format outside the GPU ingest: read by OpenCV on the host
    cv2.imwrite(bm, receipts.receipt(64)[:640, :480])
    jpgs.append(bm)
    for p in jpgs:
        res = {}
        image = cv2.imread(p)
        with contextlib.redirect_stdout(io.StringIO()):
            for crop in detector.process(image):
                _, res = recognizer.process(res, cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY))
        before = m.encoded_batches
        assert m.predict(p) == [v for k, v in res.items() if k > thresh], p
        assert (m.encoded_batches == before + 1) == (p != bm)
    m.close()
