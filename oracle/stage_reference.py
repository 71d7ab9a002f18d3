"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the reference's own Python sources for the
hot path into the git-ignored ``baseline/_ref/`` so that they travel to the GPU box
(``/root/reference`` does not exist there).

    python -m oracle.stage_reference          # run in the dev container; idempotent

Nothing is edited: every file is copied byte for byte and its SHA-256 is recorded in
``baseline/_ref/MANIFEST.json`` next to the SHA-256 of the source, so a reader can check that
what ``bench.py --impl reference`` and the drop-in tests import *is* the reference.  The files
are the ones SURVEY.md section 8 cites for the path:

    vqvae/bottleneck.py        QuantizedBottleneck (8a5-a10)
    vqvae/vqvae.py             VQVAE.encode / decode_code (8a4, a11)
    vqvae/encoder_decoder.py   the conv stacks both arms share
    utils/datasets/*.py        CodeRow / LMDBDataset / label encoders (8f N1: the reader of our rows)
    utils/misc.py, utils/distributed.py   the helper factory and rank helpers (8a1, 8e)

``baseline/_ref`` is listed in .gitignore (never in .gpurunignore); reference sources never
enter this repository's history.  The third-party packages the reference imports and this
image lacks (GANsynth_pytorch, fastai, discretization) are NOT staged -- ``oracle/ref_loader``
stubs their module-level names, exactly as it does against ``/root/reference``.
"""
import hashlib
import json
import pathlib
import shutil

ROOT = pathlib.Path(__file__).resolve().parent.parent
SOURCE = pathlib.Path("/root/reference")
TARGET = ROOT / "baseline" / "_ref"
PACKAGE = "interactive_spectrogram_inpainting"
FILES = [
    "__init__.py",
    "vqvae/__init__.py",
    "vqvae/bottleneck.py",
    "vqvae/vqvae.py",
    "vqvae/encoder_decoder.py",
    "utils/__init__.py",
    "utils/misc.py",
    "utils/distributed.py",
    "utils/datasets/__init__.py",
    "utils/datasets/label_encoders.py",
    "utils/datasets/lmdb_dataset.py",
]


def _sha(path: pathlib.Path) -> str:
    return hashlib.sha256(path.read_bytes()).hexdigest()


def staged() -> bool:
    return (TARGET / "MANIFEST.json").is_file()


def stage(force: bool = False) -> bool:
    """Copy the files; returns False (and does nothing) when the reference checkout is absent."""
    src_pkg = SOURCE / PACKAGE
    if not src_pkg.is_dir():
        return False
    manifest = {"source": str(SOURCE), "files": {}}
    for rel in FILES:
        src, dst = src_pkg / rel, TARGET / PACKAGE / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        if force or not dst.exists() or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest["files"][f"{PACKAGE}/{rel}"] = {"sha256": _sha(dst), "source_sha256": _sha(src)}
    (TARGET / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    return True


def verify() -> bool:
    """True when every staged file still hashes to what the manifest recorded."""
    if not staged():
        return False
    manifest = json.loads((TARGET / "MANIFEST.json").read_text())
    return all((TARGET / rel).is_file() and _sha(TARGET / rel) == meta["sha256"] == meta["source_sha256"]
               for rel, meta in manifest["files"].items())


if __name__ == "__main__":
    print("staged" if stage() else f"{SOURCE} not present: nothing staged", "-> verify:", verify())
