"""TEST INFRASTRUCTURE - runs the UNMODIFIED reference (`ocr/pipeline.py`, `ocr/net.py`, `ocr/server.py`) from a copy.

The reference is plain Python without a build step.  `install()` copies `/root/reference/ocr` (sources only) to the
git-ignored `baseline/_ref/ocr`, which travels to the GPU box with the repository snapshot like the built .so files
do; nothing under the repository's history holds reference sources.  `stage()` makes a scratch copy per recognition
head (the reference derives its checkpoint directory from its own location, ocr/net.py:19, and needs a writable
`save_models/` and `test/`), writes the synthetic checkpoints there and sets `prediction` / `num_classes` in the copy's
config.yml.  `imported()` puts the staged directory on sys.path - optionally behind `lightly_ocr_b200/dropin`, so that
the reference's `from net import CRAFT, CRNN` (ocr/pipeline.py:9) resolves to the CUDA drop-in while pipeline.py and
server.py themselves are the reference's files, byte for byte.

Shims, none of them touching reference files:
  * torchvision.models.vgg.model_urls   (removed from torchvision >= 0.13; modules/vgg_bn.py:6,37 only rewrites a URL)
  * stub modules lmdb, skimage, skimage.io  (imported by tools/dataset.py:7, tools/imgproc.py:3; unused on the path)
  * stub flask / werkzeug for server.py (not installed in this image): a Flask whose `route` keeps the function, a
    `jsonify` that returns the mapping, a module-level `request` the caller fills in
"""
import contextlib
import hashlib
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, "baseline", "_ref", "ocr")
UPSTREAM = "/root/reference/ocr"
DROPIN = os.path.join(ROOT, "lightly_ocr_b200", "dropin")
_SKIP = shutil.ignore_patterns("__pycache__", "*.pyc", "noteboooks", "save_models", "Dockerfile", "*.ipynb")
_REF_MODULES = ("net", "model", "modules", "tools", "pipeline", "server", "torch2onnx")


def install(force=False):
    """Copy the reference's `ocr/` package next to the repository (git-ignored).  Returns the path or None when
    /root/reference is not mounted (GPU box: the copy made in the build container is used)."""
    if not os.path.isdir(UPSTREAM):
        return INSTALLED if os.path.isdir(INSTALLED) else None
    if os.path.isdir(INSTALLED) and not force and _tree_digest(INSTALLED) == _tree_digest(UPSTREAM):
        return INSTALLED
    if os.path.isdir(INSTALLED):
        shutil.rmtree(INSTALLED)
    os.makedirs(os.path.dirname(INSTALLED), exist_ok=True)
    shutil.copytree(UPSTREAM, INSTALLED, ignore=_SKIP)
    return INSTALLED


def _tree_digest(root):
    h = hashlib.sha256()
    for d, dirs, files in sorted(os.walk(root)):
        dirs[:] = sorted(x for x in dirs if x not in ("__pycache__", "noteboooks", "save_models"))
        for f in sorted(files):
            if f.endswith((".py", ".yml")):
                h.update(os.path.relpath(os.path.join(d, f), root).encode())
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def source():
    """Directory holding the unmodified reference `ocr/` package, or None."""
    if os.path.isdir(INSTALLED):
        return INSTALLED
    if os.path.isdir(UPSTREAM):
        return UPSTREAM
    return None


def stage(head, craft_sd, crnn_sd, scratch):
    """Scratch copy of the reference configured for `head` with the given state dicts as its checkpoints."""
    import torch
    import yaml
    src = source()
    if src is None:
        raise RuntimeError("reference not installed: run oracle/ref_env.install() where /root/reference is mounted")
    dst = os.path.join(str(scratch), "ocr_" + head)
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=_SKIP)
    os.makedirs(os.path.join(dst, "test"), exist_ok=True)
    os.makedirs(os.path.join(dst, "save_models"), exist_ok=True)
    cfg_path = os.path.join(dst, "config.yml")
    with open(cfg_path) as f:
        cfg = yaml.safe_load(f)
    cfg["prediction"] = head
    cfg["num_classes"] = 37 if head == "CTC" else 38
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    torch.save(craft_sd, os.path.join(dst, "save_models", "CRAFT.pth"))
    torch.save(crnn_sd, os.path.join(dst, "save_models", "CRNN.pth"))
    return dst


def _shims(flask=False):
    import torchvision.models.vgg as tv_vgg
    if not hasattr(tv_vgg, "model_urls"):
        tv_vgg.model_urls = {"vgg16_bn": "https://download.pytorch.org/models/vgg16_bn-6c64b313.pth"}
    for name in ("lmdb", "skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    if flask and "flask" not in sys.modules:
        fl = types.ModuleType("flask")

        class Flask:
            def __init__(self, name):
                self.config = _Cfg()
                self.routes = {}

            def route(self, rule, methods=None):
                def deco(fn):
                    self.routes[rule] = fn
                    return fn
                return deco

            def run(self, **kw):
                raise RuntimeError("stub flask cannot serve")

        class _Cfg(dict):
            def from_mapping(self, **kw):
                self.update(kw)

        fl.Flask = Flask
        fl.jsonify = lambda obj: obj
        fl.request = types.SimpleNamespace(file={})
        sys.modules["flask"] = fl
        wz = types.ModuleType("werkzeug")
        wzu = types.ModuleType("werkzeug.utils")
        wzu.secure_filename = lambda s: os.path.basename(s).replace(" ", "_")
        wz.utils = wzu
        sys.modules["werkzeug"] = wz
        sys.modules["werkzeug.utils"] = wzu


def _purge():
    for m in [k for k in sys.modules if k.split(".")[0] in _REF_MODULES]:
        del sys.modules[m]


@contextlib.contextmanager
def imported(staged_dir, dropin=False, flask=False):
    """Context in which `import pipeline` (and `net`, `tools`, `server`) load from `staged_dir`.  With dropin=True the
    CUDA drop-in's `net` module shadows the reference's ocr/net.py; everything else is the reference's own file."""
    _shims(flask)
    _purge()
    paths = ([DROPIN] if dropin else []) + [staged_dir]
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    for p in reversed(paths):
        sys.path.insert(0, p)
    cwd = os.getcwd()
    # MODEL_PATH, the results file and the upload folder are relative to the cwd (net.py:19, pipeline.py:81,
    # server.py:10); from the parent directory all three resolve inside the staged copy
    os.chdir(os.path.dirname(os.path.abspath(staged_dir)))
    old_env = os.environ.get("LOCR_OCR_DIR")
    if dropin:
        os.environ["LOCR_OCR_DIR"] = staged_dir
        import importlib
        import lightly_ocr_b200.net as _n
        for e in getattr(_n, "_ENGINES", {}).values():
            e.close()
        importlib.reload(_n)
    try:
        yield staged_dir
    finally:
        os.chdir(cwd)
        for p in paths:
            if p in sys.path:
                sys.path.remove(p)
        _purge()
        if dropin:
            if old_env is None:
                os.environ.pop("LOCR_OCR_DIR", None)
            else:
                os.environ["LOCR_OCR_DIR"] = old_env


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
