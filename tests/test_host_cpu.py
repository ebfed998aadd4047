"""CPU-only checks: the C-ABI library builds, loads and exports every symbol include/locr.h declares; host logic
(comparator, converters, sharding over a 2-process gloo group) behaves like the reference's."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from lightly_ocr_b200 import build
    lib = ctypes.CDLL(build.build())
    header = open(os.path.join(ROOT, "include", "locr.h")).read()
    names = sorted(set(re.findall(r"LOCR_API[^;(]*?\b(locr_\w+)\s*\(", header)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "include/locr.h declares %s but liblocr.so does not export it" % n
    lib.locr_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.locr_version()


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lightly_ocr_b200 import bridge
    with pytest.raises(bridge.LocrError):
        bridge.Engine()


def test_compare_rects_and_sort_match_oracle_restatement():
    from lightly_ocr_b200 import hostops
    from oracle import ocr_ref
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(0, 90))
        rects = []
        for _ in range(n):
            y0, x0 = int(rng.integers(0, 600)), int(rng.integers(0, 900))
            rects.append([y0, x0, y0 + int(rng.integers(1, 60)), x0 + int(rng.integers(1, 200))])
        assert hostops.sort_rects(rects) == ocr_ref.sort_rects(rects)


def test_overlapped_sort_permutation_equals_sorting_the_rects():
    """OcrRunner recognises the crops in label order while a helper thread sorts `rect + [position]` lists with the same
    comparator (it only reads elements 0..3) and permutes the outputs afterwards: the permutation must reproduce
    sorted(rects, key=cmp_to_key(compare_rects)) exactly, duplicates and degenerate rects included - also on the live
    reference's golden rects."""
    from lightly_ocr_b200 import hostops
    rng = np.random.default_rng(1)
    sets = []
    for trial in range(200):
        k = int(rng.integers(0, 120))
        ys, xs = rng.integers(0, 1200, k), rng.integers(0, 900, k)
        r = np.stack([ys, xs, ys + rng.integers(-2, 40, k), xs + rng.integers(-2, 200, k)], 1).astype(np.int32)
        if trial % 3 == 0 and k > 4:
            r[k // 2] = r[k // 3]
        sets.append(r.tolist())
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_ctc.npz"))
    sets += [g["maps1_rects"].tolist(), g["maps2_rects"].tolist()]
    for flat in sets:
        want = hostops.sort_rects([list(x) for x in flat])
        perm = [t[4] for t in hostops.sort_rects([flat[j] + [j] for j in range(len(flat))])]
        assert [flat[q] for q in perm] == want
        assert sorted(perm) == list(range(len(flat)))


def test_sorted_rects_match_reference_golden():
    from lightly_ocr_b200 import hostops
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_ctc.npz"))
    for s in (1, 2):
        assert np.array_equal(np.array(hostops.sort_rects(g["maps%d_rects" % s].tolist()), np.int32),
                              g["maps%d_sorted" % s])


def test_converters_match_reference_known_answers():
    from lightly_ocr_b200 import hostops
    conv = hostops.CTCLabelConverter("abcdefghijklmnopqrstuvwxyz")
    assert conv.decode([6, 9, 6, 1], [4]) == ["fifa"]          # reference ocr/test/utils_test.py:37-39
    assert conv.decode([5, 5, 0, 1], [4]) == ["ea"]            # :41-43
    att = hostops.AttnLabelConverter(hostops.ALPHABET)
    assert att.decode(np.array([[0, 12, 1, 2]]), [25]) == ["[GO]a[s]0"]


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from lightly_ocr_b200 import shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["MASTER_PORT"], rank=rank, world_size=world)
mine = shard.shard_indices(11, rank, world)
res = shard.gather_in_order([(i, "r%%d" %% i) for i in mine], 11, rank, world)
t = shard.max_over_ranks(1.0 + rank)
if rank == 0:
    assert res == ["r%%d" %% i for i in range(11)], res
    assert t == float(world), t
    print("OK")
dist.barrier()
dist.destroy_process_group()
""" % ROOT


def test_sharding_two_process_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29713")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]


def test_serving_front_batches_and_filters_without_gpu(tmp_path, monkeypatch):
    """Host logic of lightly_ocr_b200/serve.py (request batching, per-request fan-out, the `k > thresh` filter of
    pipeline.serveModel.predict, error isolation) with the GPU runner replaced by a stub."""
    import threading
    import cv2
    import yaml
    d = tmp_path / "ocr"
    d.mkdir()
    cfg = yaml.safe_load(open(os.path.join(ROOT, "lightly_ocr_b200", "config.yml")))
    cfg["prediction"], cfg["num_classes"] = "CTC", 37
    yaml.safe_dump(cfg, open(str(d / "config.yml"), "w"))
    monkeypatch.setenv("LOCR_OCR_DIR", str(d))
    import importlib
    import lightly_ocr_b200.net as net
    importlib.reload(net)
    import lightly_ocr_b200.serve as serve
    importlib.reload(serve)

    class StubRunner:
        def __init__(self):
            self.calls = []

        def ocr(self, images):
            # one "crop" per 100 rows of the image; text = image height, confidence alternates around the threshold
            self.calls.append(len(images))
            per_image, text, conf, eos = [], [], [], []
            for im in images:
                n = im.shape[0] // 100
                per_image.append([[0, 0, 1, 1]] * n)
                for k in range(n):
                    text.append("h%dk%d" % (im.shape[0], k))
                    conf.append(0.9 if k % 2 == 0 else 0.5)
                    eos.append(1)
            return per_image, dict(text=text, conf=np.array(conf, np.float32), has_eos=np.array(eos, np.int32))

        def ocr_encoded(self, blobs):
            # the GPU JPEG route: stand-in decodes with OpenCV and reports the route taken; like liblocr it refuses
            # a file whose entropy-coded data ends early (here: no EOI marker)
            self.encoded_calls = getattr(self, "encoded_calls", 0) + 1
            from lightly_ocr_b200 import bridge as _b
            if any(b[:2] == b"\xff\xd8" and not b.rstrip(b"\0").endswith(b"\xff\xd9") for b in blobs):
                raise _b.LocrError("liblocr error -2: JPEG: premature end of the entropy-coded data")
            per_image, out = self.ocr([cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in blobs])
            return per_image, out, [(0, 0)] * len(blobs)

        def close(self):
            pass

    def fake_load(self):
        self.head = "CTC"
        self.runner = StubRunner()
        self.detector = self.recognizer = self.runner

    monkeypatch.setattr(serve.serveModel, "loadModel", fake_load)
    m = serve.serveModel(config_file="config.yml", thresh=0.7, docker=True, max_batch=4, max_wait_ms=200)
    paths = []
    for i, h in enumerate((100, 200, 300, 400, 500, 600)):
        p = str(d / ("u%d.bmp" % i))                     # a format outside the GPU ingest: read by OpenCV on the host
        cv2.imwrite(p, np.full((h, 50, 3), 255, np.uint8))
        paths.append(p)
    got = [None] * len(paths)
    th = [threading.Thread(target=lambda i=i: got.__setitem__(i, serve.api_response(m, paths[i]))) for i in range(len(paths))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=30)
    for i, h in enumerate((100, 200, 300, 400, 500, 600)):
        body, status = got[i]
        want = [["h%dk%d" % (h, k)] for k in range(h // 100) if k % 2 == 0]     # CTC values are one-element lists
        assert status == 200 and body == {"status": "OK", "results": {k: v for k, v in enumerate(want)}}
    assert sum(m.batches) == 6 and max(m.batches) <= 4 and len(m.batches) < 6      # requests were batched
    with pytest.raises(ValueError):
        m.predict(str(d / "missing.png"))
    assert m.predict(paths[0]) == [["h100k0"]]
    # JPEG and PNG uploads stay encoded and take the GPU-decode route (runner.ocr_encoded); other formats go through OpenCV
    jp = str(d / "u.jpg")
    cv2.imwrite(jp, np.full((300, 50, 3), 255, np.uint8))
    assert m.encoded_batches == 0
    assert m.predict(jp) == [["h300k0"], ["h300k2"]]
    assert m.encoded_batches == 1 and m.runner.encoded_calls == 1
    assert m.predict(paths[2]) == [["h300k0"], ["h300k2"]] and m.encoded_batches == 1
    pg = str(d / "u.png")
    cv2.imwrite(pg, np.full((200, 50, 3), 255, np.uint8))
    assert m.predict(pg) == [["h200k0"]] and m.encoded_batches == 2
    bad = str(d / "broken.jpg")
    with open(bad, "wb") as f:
        f.write(open(jp, "rb").read()[:40])          # truncated header: neither reader accepts it
    with pytest.raises(ValueError):
        m.predict(bad)
    # a JPEG with a valid header whose SCAN is truncated, batched with good uploads: the batched GPU call fails, the
    # members are retried one at a time, the good ones are served, and the broken one falls back to OpenCV, which
    # returns a partly grey image exactly like the reference's cv2.imread (the request succeeds)
    big = str(d / "big.jpg")
    rng = np.random.default_rng(0)
    cv2.imwrite(big, rng.integers(0, 256, (300, 64, 3), dtype=np.uint8))
    cut = str(d / "cut.jpg")
    blob = open(big, "rb").read()
    with open(cut, "wb") as f:
        f.write(blob[:len(blob) * 2 // 3])
    from lightly_ocr_b200 import bridge
    assert bridge.jpeg_info(open(cut, "rb").read())[:2] == (300, 64)         # the header alone looks fine
    m.max_wait = 0.5
    before = m.retried_batches
    trio = [jp, cut, big]
    got3 = [None] * 3
    th = [threading.Thread(target=lambda i=i: got3.__setitem__(i, serve.api_response(m, trio[i]))) for i in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=30)
    assert all(g is not None and g[1] == 200 for g in got3)
    assert got3[0][0]["results"] == {0: ["h300k0"], 1: ["h300k2"]} and got3[2][0]["results"] == got3[0][0]["results"]
    assert got3[1][0]["results"] == got3[0][0]["results"]                   # OpenCV's partly grey 300-row image
    assert m.retried_batches == before + 1
    m.close()


def test_installed_reference_copy_is_unmodified():
    """baseline/_ref/ocr (git-ignored, made by oracle/ref_env.install) must be the reference's files, byte for byte:
    checked wherever both the copy and /root/reference are present (the authoring container)."""
    from oracle import ref_env
    if not (os.path.isdir(ref_env.INSTALLED) and os.path.isdir(ref_env.UPSTREAM)):
        pytest.skip("needs both baseline/_ref/ocr and /root/reference")
    assert ref_env._tree_digest(ref_env.INSTALLED) == ref_env._tree_digest(ref_env.UPSTREAM)
    for name in ("pipeline.py", "net.py", "server.py", "tools/det_utils.py"):
        with open(os.path.join(ref_env.INSTALLED, name), "rb") as a, open(os.path.join(ref_env.UPSTREAM, name), "rb") as b:
            assert a.read() == b.read(), name


def test_evaluation_loop_host_logic():
    """lightly_ocr_b200.evaluate.evaluation (the reference's validation loop, ocr/train/crnn.py:142-240) with a stand-in
    engine: targets reach the engine exactly as the reference's converters encode them, the loss is the Averager's mean
    of the batch costs, the accuracy counts the engine's flags, `max_iter` bounds the loop, and the returned preds_ /
    confidence_ / label are those of the last batch (the reference overwrites them per batch)."""
    from lightly_ocr_b200 import evaluate

    class Fake:
        head = 0          # bridge.HEAD_CTC

        def __init__(self):
            self.calls = []

        def evaluate(self, crops, targets, target_len):
            self.calls.append((len(crops), np.array(targets), np.array(target_len)))
            n = len(crops)
            return dict(cost=0.5 * len(self.calls), loss=np.ones(n, np.float32), correct=np.array([1] + [0] * (n - 1), np.int32),
                        ids=np.zeros((n, 26), np.int32), text=["t%d" % i for i in range(n)], conf=np.linspace(0.1, 0.9, n))

    eng = Fake()
    batches = [([np.zeros((8, 8), np.uint8)] * 3, ["ab", "c", "0z9"]), ([np.zeros((8, 8), np.uint8)] * 2, ["q", "rs"]),
               ([np.zeros((8, 8), np.uint8)], ["never"])]
    loss, acc, preds_, conf_, label, infer_, n = evaluate.evaluation(eng, batches, {"batch_max_len": 25, "max_iter": 2})
    assert len(eng.calls) == 2 and n == 5
    assert abs(loss - (0.5 + 1.0) / 2) < 1e-9 and abs(acc - 2 / 5 * 100) < 1e-9
    assert preds_ == ["t0", "t1"] and label == ["q", "rs"] and len(conf_) == 2 and infer_ >= 0
    # CTC targets: blank = 0, '0' -> 1 ... 'z' -> 36, concatenated, with the lengths
    assert eng.calls[0][1].tolist() == [11, 12, 13, 1, 36, 10] and eng.calls[0][2].tolist() == [2, 1, 3]
    eng2 = Fake()
    eng2.head = 1         # attention: [GO] = 0, [s] = 1, rows of batch_max_len + 2
    evaluate.evaluation(eng2, batches[:1], {"batch_max_len": 25})
    tg = eng2.calls[0][1]
    assert tg.shape == (3, 27) and tg[0, :4].tolist() == [0, 12, 13, 1] and tg[2, :5].tolist() == [0, 2, 37, 11, 1]
    assert eng2.calls[0][2].tolist() == [3, 2, 4]



def test_confidence_line_formats_like_the_reference_tensor():
    """net.CRNN.process prints `confidence score: {confidence:.4f}` of a 0-d fp32 tensor (reference ocr/net.py:192); the
    drop-in formats the Python float it built the tensor from - the same characters for every fp32 value."""
    import torch
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.random(5000).astype(np.float32),
                           np.float32([0, 1, 0.99995, 0.00005, 1e-8, 0.12345, 0.5, 0.99999994, 0.00015, 0.99985])])
    for v in vals:
        conf = float(v)
        assert f"{torch.tensor(conf, dtype=torch.float32):.4f}" == f"{conf:.4f}"
