"""Host logic of bench.py without a GPU: the legs, their bookkeeping and the JSON line (keys the driver reads) are run
against a stand-in for bridge.OcrRunner.  No number printed here means anything - the point is that a typo in the bench
cannot cost the round its headline line."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Runner:
    """what bench.py touches of bridge.OcrRunner"""
    made = []

    def __init__(self, device_id=0, act_dtype=None, head="CTC", precision=0):
        self.head, self.precision, self.n, self.closed = head, precision, 0, False
        _Runner.made.append(self)

    def load_state_dict(self, model, sd):
        assert sd["head"] in (None, self.head)

    def _out(self, n):
        self.n += 7 * n
        return [[(0, 0, 4, 4)] * 3 for _ in range(n)], {"text": ["x"] * (3 * n)}

    def ocr(self, images):
        assert not self.closed
        return self._out(len(images))

    def ocr_resident(self, n):
        return self._out(n)

    def launch_count(self):
        return self.n

    def timer_start(self):
        pass

    def timer_stop(self):
        return 12.5

    def profile(self, on):
        pass

    def profile_read(self):
        return 10.0, 1e12, 40

    def profile_layers(self):
        return [("conv", 10.0, 40), ("other", 2.0, 9)]

    def close(self):
        self.closed = True


@pytest.mark.parametrize("argv", [[], ["--head", "Attention"], ["--no-other-head", "--no-other-precision"]])
def test_bench_line_shape_with_a_stand_in_engine(monkeypatch, argv):
    import torch
    import bench
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import weights
    _Runner.made = []
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(bridge, "OcrRunner", _Runner)
    monkeypatch.setattr(weights, "craft_calibrated", lambda *a, **k: {"head": None})
    monkeypatch.setattr(weights, "crnn_calibrated", lambda seed, head: {"head": head})
    monkeypatch.setattr(bench, "make_receipts", lambda rank, n: [np.zeros((8, 8, 3), np.uint8) for _ in range(n)])
    monkeypatch.setenv("LOCR_BENCH_SAMPLER", "0")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "3", "--no-dropin", "--no-cpu-baseline"] + argv)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.main()
    line = json.loads(out.getvalue().strip().splitlines()[-1])
    head = "Attention" if "Attention" in argv else "CTC"
    assert line["metric"] == bench.metric_name(head) and line["unit"] == "receipts/s" and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["warmup"] == 3 and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == bench.PASSES * bench.PER_PASS * 8 * 8 * 3 and line["e2e"]["d2h_bytes_per_step"] > 0
    r = line["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["kernel_share_of_gpu_time"] - 10.0 / 12.0) < 1e-12
    assert "workload" in line["config"] and "clocks" in line
    if "--no-other-head" in argv:
        assert "other_head" not in line and "other_precision" not in line
    else:
        oh = line["other_head"]
        other = "CTC" if head == "Attention" else "Attention"
        assert oh["head"] == other and oh["metric"] == bench.metric_name(other) and oh["value"] > 0 and oh["n_gpus"] == 1
        assert ("config %d" % (5 if other == "Attention" else 4)) in oh["workload"]
        assert line["other_precision"]["precision"] == "exact" and line["other_precision"]["head"] == head
        # main legs, exact leg, other-head leg: LANES runners each, every one closed again
        assert [m.head for m in _Runner.made] == [head] * (2 * bench.LANES) + [other] * bench.LANES
    assert all(m.closed for m in _Runner.made)


_WORKER = r"""
import io, json, os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import bench
from lightly_ocr_b200 import bridge
from lightly_ocr_b200.synth import weights
from test_bench_cpu import _Runner
bridge.OcrRunner = _Runner
weights.craft_calibrated = lambda *a, **k: {"head": None}
weights.crnn_calibrated = lambda seed, head: {"head": head}
bench.make_receipts = lambda rank, n: [np.zeros((8, 8, 3), np.uint8) for _ in range(n)]
sys.argv = ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3"]
bench.main()
"""


def test_bench_two_ranks_gloo(tmp_path):
    """The launch the driver uses for N > 1 (one process per GPU, RANK / WORLD_SIZE / MASTER_* from the environment), two
    ranks over gloo with the stand-in engine: every leg's barriers and reductions line up on both ranks, rank 0 alone
    prints the one JSON line, whole-job values count both ranks, per-rank times are reported."""
    import subprocess
    script = tmp_path / "w.py"
    script.write_text(_WORKER % (ROOT, ROOT))
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29719", LOCR_BENCH_BACKEND="gloo",
               LOCR_BENCH_SAMPLER="0")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert outs[1][0].strip() == "", "rank 1 printed: %r" % outs[1][0]
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["value"] > 0 and line["e2e"]["value"] > 0
    assert len(line["ranks"]["per_rank_dev_s"]) == 2 and len(line["e2e"]["ranks"]["per_rank_wall_s"]) == 2
    # the stand-in reports 3 crops per receipt on every rank: whole-job crops / receipts = 3
    assert abs(line["crops_per_sec"] / line["value"] - 3.0) < 1e-9
    assert line["other_head"]["n_gpus"] == 2 and line["other_head"]["head"] == "Attention" and line["other_head"]["value"] > 0
    assert abs(line["other_head"]["crops_per_sec"] / line["other_head"]["value"] - 3.0) < 1e-9
    assert line["other_precision"]["precision"] == "exact"
    assert "e2e_dropin" not in line and "cpu_baseline" not in line     # rank 0 at N = 1 only
