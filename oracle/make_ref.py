#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY - recipe that stages the UNMODIFIED reference for the GPU box.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python (its arithmetic lives in PyTorch ATen); there is nothing to compile. `/root/reference`
does not exist on the GPU box, so this script copies the files of the hot path and its driver - byte for byte, no
edits - into `oracle/_ref/`:

    main.py  util.py  dataset.py  models/graphcnn.py  models/mlp.py  models/discriminator.py

`oracle/_ref/` is listed in .gitignore (reference sources never enter this repository's history) but NOT in
.gpurunignore, so the copy travels to the box with the snapshot, like the built libgnm.so. `MANIFEST.json` records the
sha256 of every copied file; `oracle/ref_arm.py` refuses a copy whose hashes do not match its manifest.

Used by: bench.py (`--impl reference` and the `cpu_baseline` leg: the reference's own `GIN_InfoMaxReg` on the host
cores, kind "reference") and tests/test_reference_driver.py (the reference's own main.py train()/test()/
get_saliency_map()/get_latent_space() executed against the repo's `models/` shim). The product never imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GNM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["main.py", "util.py", "dataset.py", "models/graphcnn.py", "models/mlp.py", "models/discriminator.py"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("oracle/make_ref.py: %s not present (GPU box?) - keeping whatever oracle/_ref holds" % REF)
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
        assert manifest[rel] == sha256(src)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/make_ref.py: staged %d reference files into %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
