"""TEST / BASELINE INFRASTRUCTURE ONLY - loads the UNMODIFIED reference staged in `oracle/_ref/` (oracle/make_ref.py:
one archive of the six reference files + a sha256 manifest; unpacked here into a per-process temporary directory).

Only tests/, __graft_entry__.smoke() and bench.py's `cpu_baseline` / `--impl reference` legs may import this module;
the product (`graph_neural_mapping_b200/`, `models/`) never does.

Two things are offered:
  * `reference_model_class()`  - the reference's own `GIN_InfoMaxReg` (models/graphcnn.py:12), loaded under a private
    module name so it cannot collide with the repo's `models/` shim. bench.py times it on the host cores as the CPU
    arm (kind "reference").
  * `reference_main(models_from)` - the reference's own `main.py` module (train / test / pass_data_iteratively /
    get_saliency_map / get_latent_space, main.py:19-96), imported with `models` resolving either to the reference's
    classes ("reference") or to the repo-root `models/` shim ("repo") - the drop-in claim of SURVEY 8(b), executed.
"""
import atexit
import hashlib
import importlib
import importlib.util
import json
import os
import shutil
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE_DIR = os.path.join(HERE, "_ref")
REPO = os.path.dirname(HERE)
REF_DIR = None          # where the archive is unpacked (set by verify())


def available():
    if not os.path.isfile(os.path.join(STAGE_DIR, "MANIFEST.json")):
        return False
    with open(os.path.join(STAGE_DIR, "MANIFEST.json")) as f:
        return os.path.isfile(os.path.join(STAGE_DIR, json.load(f).get("archive", "")))


def verify():
    """Unpack the staged archive (once per process) and check that every file is byte-identical to what make_ref.py
    read from /root/reference."""
    global REF_DIR
    with open(os.path.join(STAGE_DIR, "MANIFEST.json")) as f:
        meta = json.load(f)
    manifest = meta["files"]
    if REF_DIR is None:
        d = tempfile.mkdtemp(prefix="gnm_ref_")
        atexit.register(shutil.rmtree, d, True)
        with tarfile.open(os.path.join(STAGE_DIR, meta["archive"]), "r:gz") as tar:
            for m in tar.getmembers():
                if m.name not in manifest or not m.isfile():
                    raise RuntimeError("unexpected member %r in the staged reference archive" % m.name)
            tar.extractall(d)
        REF_DIR = d
    for rel, want in manifest.items():
        with open(os.path.join(REF_DIR, rel), "rb") as f:
            got = hashlib.sha256(f.read()).hexdigest()
        if got != want:
            raise RuntimeError("staged reference file %s does not match its manifest: the reference copy was edited" % rel)
    return sorted(manifest)


def _purge(names):
    for m in [m for m in sys.modules if m in names or any(m.startswith(n + ".") for n in names)]:
        del sys.modules[m]


def reference_model_class():
    """The reference's GIN_InfoMaxReg. graphcnn.py:6-9 appends "models/" (relative to the cwd) to sys.path and imports
    top-level `mlp` / `discriminator`; both are pre-loaded here from oracle/_ref/models so the cwd does not matter."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    verify()
    mods = {}
    saved = {k: sys.modules.get(k) for k in ("mlp", "discriminator")}
    try:
        for name in ("mlp", "discriminator"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, "models", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
        spec = importlib.util.spec_from_file_location("gnm_reference_graphcnn", os.path.join(REF_DIR, "models", "graphcnn.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref.GIN_InfoMaxReg


def reference_main(models_from):
    """Import oracle/_ref/main.py as a module. models_from = "repo": `from models.graphcnn import *` (main.py:9)
    resolves to the repo-root `models/` package (it has an __init__.py, so it wins over the reference's namespace
    directory) - the reference driver then runs UNCHANGED on the libgnm kernels. "reference": resolves to the
    reference's own classes."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    verify()
    assert models_from in ("repo", "reference")
    _purge({"models", "util", "dataset", "mlp", "discriminator"})
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    try:
        if models_from == "repo":
            sys.path[:0] = [REPO, REF_DIR]
        else:
            sys.path[:0] = [REF_DIR]
            sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
            os.chdir(REF_DIR)                      # graphcnn.py:7 appends the cwd-relative "models/"
        spec = importlib.util.spec_from_file_location("gnm_reference_main_" + models_from, os.path.join(REF_DIR, "main.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        where = sys.modules["models.graphcnn"].__file__
        inside = os.path.abspath(where).startswith(REF_DIR)
        if inside != (models_from == "reference"):
            raise RuntimeError("main.py resolved models.graphcnn to %s (wanted the %s classes)" % (where, models_from))
    finally:
        sys.path[:] = [p for p in saved_path]
        os.chdir(saved_cwd)
        _purge({"models", "util", "dataset", "mlp", "discriminator"})
    return mod
