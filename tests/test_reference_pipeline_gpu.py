"""The reference's OWN `ocr/pipeline.py` and `ocr/server.py`, unmodified (copied by oracle/ref_env.py into the
git-ignored baseline/_ref/ocr), executed against the CUDA drop-in: `lightly_ocr_b200/dropin` sits ahead of the
reference directory on sys.path, so `from net import CRAFT, CRNN` (ocr/pipeline.py:9) binds the B200 classes while
`prepModel`, `getText` and `serveModel` (pipeline.py:47-112) and `parseText` (server.py:49-53) are the reference's code.

Compared with the goldens that oracle/make_golden.py recorded from the LIVE reference (its own net.py on the CPU) on
the same receipts and checkpoints: same number of results, >= 99.5 % identical strings.
"""
import contextlib
import io
import os

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _need_reference():
    from oracle import ref_env
    if ref_env.source() is None:
        pytest.skip("baseline/_ref/ocr not installed (python -m oracle.ref_env where /root/reference is mounted)")
    return ref_env


def _values(res):
    return [v[0] if isinstance(v, list) else v for v in res.values()]


@pytest.mark.parametrize("head", ["CTC", "Attention"])
def test_unmodified_pipeline_against_dropin(tmp_path, head):
    ref_env = _need_reference()
    from oracle import receipts, weights
    gz = np.load(os.path.join(GOLDEN, "ref_%s.npz" % head.lower()))
    dst = ref_env.stage(head, weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, head), tmp_path)
    with ref_env.imported(dst, dropin=True):
        import net
        import pipeline
        assert os.path.samefile(os.path.dirname(pipeline.__file__), dst)
        assert os.path.samefile(os.path.dirname(net.__file__), ref_env.DROPIN)
        with open(pipeline.__file__, "rb") as a, open(os.path.join(ref_env.source(), "pipeline.py"), "rb") as b:
            assert a.read() == b.read()                       # the reference's file, byte for byte
        detector, recognizer = pipeline.prepModel(pipeline.CONFIG)
        assert type(detector).__module__.startswith("lightly_ocr_b200")
        got, counts = [], []
        for rid in gz["e2e_receipts"]:
            path = os.path.join(str(tmp_path), "receipt%d.png" % int(rid))
            cv2.imwrite(path, receipts.receipt(int(rid)))
            out = io.StringIO()
            with contextlib.redirect_stdout(out):
                res = pipeline.getText(path, detector, recognizer)       # write=True: test/results.txt as well
            vals = _values(res)
            got.extend(vals)
            counts.append(len(vals))
            assert "wrote results to" in out.getvalue()
            with open(os.path.join(dst, "test", "results.txt")) as f:
                assert len(f.read().splitlines()) == len(res)
        want = [str(t) for t in gz["e2e_text"]]
        assert counts == [int(c) for c in gz["e2e_counts"]]
        same = sum(g == w for g, w in zip(got, want))
        print("%s: unmodified pipeline.getText over the drop-in: %d / %d strings identical to the live reference" %
              (head, same, len(want)))
        assert same / len(want) >= 0.995
        # serveModel.predict (pipeline.py:89-112): results above the confidence threshold, in order
        m = pipeline.serveModel("config.yml", 0.7, False)
        rid = int(gz["e2e_receipts"][0])
        with contextlib.redirect_stdout(io.StringIO()):
            pred = m.predict(os.path.join(str(tmp_path), "receipt%d.png" % rid))
        n0 = int(gz["e2e_counts"][0])
        gold = [(str(t), float(c)) for t, c in zip(gz["e2e_text"][:n0], gz["e2e_conf"][:n0])]
        # confidences within 2 % of the threshold may fall on either side of it (fp16 storage vs fp32)
        sure = [t for t, c in gold if c > 0.7 * 1.02]
        maybe = [t for t, c in gold if c > 0.7 * 0.98]
        flat = [p[0] if isinstance(p, list) else p for p in pred]
        assert len(sure) <= len(flat) <= len(maybe)
        it = iter(flat)
        assert all(any(s == g for g in it) for s in sure)              # `sure` is a subsequence of the prediction


def test_unmodified_server_route_against_dropin(tmp_path):
    """server.py's POST /api handler (server.py:49-53) with a stub flask (flask is not installed in this image): the
    upload is saved by getPath (server.py:24-40) and recognised by serveModel.predict."""
    ref_env = _need_reference()
    from oracle import receipts, weights
    dst = ref_env.stage("CTC", weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, "CTC"), tmp_path)
    with ref_env.imported(dst, dropin=True, flask=True):
        import flask
        import pipeline
        import server
        server.m = pipeline.serveModel("config.yml", 0.7, False)
        ok, buf = cv2.imencode(".png", np.ascontiguousarray(receipts.receipt(1)[:640, :480]))
        assert ok

        class Upload:
            filename = "receipt one.png"

            def save(self, path):
                with open(path, "wb") as f:
                    f.write(buf.tobytes())

            def __bool__(self):
                return True

        server.request.file = {"file": Upload()}
        assert server.request is flask.request
        with contextlib.redirect_stdout(io.StringIO()):
            body, code = server.parseText()
        assert code == 200 and body["status"] == "OK"
        assert len(body["results"]) > 5 and all(isinstance(k, int) for k in body["results"])
        assert os.path.exists(os.path.join(dst, "test", "receipt_one.png"))
        online, code = server.isOnline()
        assert code == 200 and online == {"status": "online"}
