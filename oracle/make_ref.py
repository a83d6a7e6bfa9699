#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY - recipe that stages the UNMODIFIED reference for the GPU box.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python (its arithmetic lives in PyTorch ATen); there is nothing to compile. `/root/reference`
does not exist on the GPU box, so this script packs the files of the hot path and its driver - byte for byte, no
edits - into ONE archive, `oracle/_ref/reference_sources.tar.gz`:

    main.py  util.py  dataset.py  models/graphcnn.py  models/mlp.py  models/discriminator.py

`oracle/_ref/` is listed in .gitignore (reference sources never enter this repository's history or its source tree as
files) but NOT in .gpurunignore, so the archive travels to the box with the snapshot, like the built libgnm.so.
`MANIFEST.json` records the sha256 of every packed file; `oracle/ref_arm.py` unpacks the archive into a per-process
temporary directory and refuses files whose hashes do not match the manifest.

Used by: bench.py (`--impl reference` and the `cpu_baseline` leg: the reference's own `GIN_InfoMaxReg` on the host
cores, kind "reference") and tests/test_reference_driver.py (the reference's own main.py train()/test()/
get_saliency_map()/get_latent_space() executed against the repo's `models/` shim). The product never imports it.
"""
import hashlib
import io
import json
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GNM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["main.py", "util.py", "dataset.py", "models/graphcnn.py", "models/mlp.py", "models/discriminator.py"]
ARCHIVE = "reference_sources.tar.gz"


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
    os.makedirs(DST, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for rel in FILES:
            src = os.path.join(REF, rel)
            manifest[rel] = sha256(src)
            info = tar.gettarinfo(src, arcname=rel)
            info.mtime = 0                      # reproducible archive
            with open(src, "rb") as f:
                tar.addfile(info, f)
    with open(os.path.join(DST, ARCHIVE), "wb") as f:
        f.write(buf.getvalue())
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "archive": ARCHIVE, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/make_ref.py: packed %d reference files into %s" % (len(FILES), os.path.join(DST, ARCHIVE)))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
