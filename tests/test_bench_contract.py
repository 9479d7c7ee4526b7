"""bench.py's reference arm runs anywhere (it times the reference's CPU path: the unmodified model files under oracle/_ref when
oracle/make_ref.py has put them there, else the oracle port): the JSON line it prints must carry the keys the driver reads.
(The GPU arm's line is checked on the GPU box by the driver itself.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoder_train_captions_per_sec" and d["unit"] == "captions/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "knowledge_aware", "models.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_oracle_ref_recipe_copies_the_reference_models_byte_for_byte(tmp_path):
    """oracle/make_ref.py (run by __graft_entry__.build() where the reference checkout exists): the files under oracle/_ref must be
    byte-identical to the reference's, the manifest must record their hashes, and the loader must give working CPU modules."""
    import hashlib

    import pytest

    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout on this machine")
    sys.path.insert(0, ROOT)
    from oracle import make_ref, ref_loader

    m = make_ref.make(ref)
    for rel, info in m["files"].items():
        got = open(os.path.join(ROOT, "oracle", "_ref", rel), "rb").read()
        assert got == open(os.path.join(ref, info["from"]), "rb").read()
        assert hashlib.sha256(got).hexdigest() == info["sha256"]
    assert ref_loader.available()
    mod = ref_loader.load_models("G")
    assert str(mod.device) == "cpu" and hasattr(mod, "DecoderTransformer")
    # and the checkout stays out of the repository history
    assert "oracle/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()
