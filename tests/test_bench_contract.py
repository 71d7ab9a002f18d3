"""bench.py's reference arm runs on the CPU (the unmodified reference VQVAE.encode behind the
front-end restatement when the reference is present or staged, else the oracle port): check
the JSON line it prints against the driver's contract.  The B200 arm needs a GPU and is exercised on the box."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "notes/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("notes/sec coded") and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    base = line["cpu_baseline"]
    from oracle import ref_loader
    assert base["kind"] == ("reference" if ref_loader.available() else "port") and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == "notes/s"
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
