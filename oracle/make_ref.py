"""
Recipe for oracle/_ref/ — TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product package).

The reference is pure Python: there is nothing to compile, but its three model files can travel to the GPU box as they are.
This script copies them UNMODIFIED, byte for byte, from the read-only reference checkout into oracle/_ref/ (git-ignored, so
the repository history stays free of reference sources; not gpurun-ignored, so the files reach the GPU box like a built .so)
and records their SHA-256 in oracle/_ref/MANIFEST.json:

    /root/reference/geo-aware/models.py             -> oracle/_ref/geo_aware/models.py
    /root/reference/knowledge-aware/models.py        -> oracle/_ref/knowledge_aware/models.py
    /root/reference/news-knowledge-aware/models.py   -> oracle/_ref/news_knowledge_aware/models.py

    python oracle/make_ref.py [--reference /root/reference]

`bench.py --impl reference` and bench.py's `cpu_baseline` legs load these modules through oracle/ref_loader.py and time the
reference's own CPU implementation (kind "reference"); without oracle/_ref they fall back to the oracle port (kind "port").
__graft_entry__.build() runs this recipe whenever the reference checkout is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
VARIANTS = {"geo_aware": "geo-aware", "knowledge_aware": "knowledge-aware", "news_knowledge_aware": "news-knowledge-aware"}


def make(reference: str = "/root/reference") -> dict:
    manifest = {"reference": reference, "files": {}}
    for pkg, sub in VARIANTS.items():
        src = os.path.join(reference, sub, "models.py")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        dst_dir = os.path.join(OUT, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        dst = os.path.join(dst_dir, "models.py")
        shutil.copyfile(src, dst)
        manifest["files"][f"{pkg}/models.py"] = {"from": f"{sub}/models.py", "sha256": hashlib.sha256(open(dst, "rb").read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return manifest


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--reference") + 1] if "--reference" in sys.argv else "/root/reference"
    m = make(ref)
    for k, v in m["files"].items():
        print(f"oracle/_ref/{k}  <-  {v['from']}  sha256 {v['sha256'][:16]}")
