"""Recipe that stages the UNMODIFIED reference under oracle/_ref/ so that the CPU legs of bench.py (`--impl reference`,
`cpu_baseline`) can time the reference's OWN code on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py            # run in the build container (also run by __graft_entry__.build())

The reference is pure Python with no build metadata (nothing to `pip install`): the product files are copied byte for
byte -- quantization/, models/, videosets/, methods/, utils.py, configs/ -- and oracle/_ref/MANIFEST.json records the
sha256 of every file next to the sha256 of its source.  oracle/_ref/ is git-ignored (reference sources never enter this
repository's history) but not gpurun-ignored, so it travels to the GPU box with the snapshot.  The three packages the
reference imports that this image lacks (timm, pytorch_msssim, hadamard_transform) come from oracle/ref_shims at run time.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NQ_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
ITEMS = ["quantization", "models", "videosets", "methods", "configs", "utils.py"]
KEEP = (".py", ".yaml")


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage() -> bool:
    if not os.path.isdir(SRC):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for item in ITEMS:
        s = os.path.join(SRC, item)
        if os.path.isfile(s):
            files = [(s, os.path.join(DST, item))]
        else:
            files = []
            for root, _, names in os.walk(s):
                for n in sorted(names):
                    if n.endswith(KEEP):
                        p = os.path.join(root, n)
                        files.append((p, os.path.join(DST, os.path.relpath(p, SRC))))
        for a, b in files:
            os.makedirs(os.path.dirname(b), exist_ok=True)
            shutil.copyfile(a, b)
            manifest[os.path.relpath(b, DST)] = {"sha256": sha(b), "source_sha256": sha(a)}
            assert manifest[os.path.relpath(b, DST)]["sha256"] == manifest[os.path.relpath(b, DST)]["source_sha256"]
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage()
    print(f"oracle/_ref: {'staged from ' + SRC if ok else 'reference not present at ' + SRC + ' (nothing done)'}")
    sys.exit(0)
